// Drop-in side of the reference's include/ORBmatcher.h (class ORB_SLAM2::ORBmatcher,
// /root/reference/include/ORBmatcher.h:37-102) over the C ABI of orb_b200.h.
//
// The reference's 11 searches walk Frame / KeyFrame / MapPoint pointer graphs that are out of scope here
// (SURVEY.md section 2), so the searches are TEMPLATES over those types: they compile against the
// reference's own Frame.h / KeyFrame.h when instantiated there, and against light stand-ins in
// tests/cpp. What they do is always the same split (hard part 6):
//   host   : candidate gating in the reference's order (GetFeaturesInArea, src/Frame.cc:327-380)
//   device : every Hamming distance of the gated (query, candidate) pairs in one launch
//   host   : the reference's own ordered best/second/ratio/one-to-one logic, replayed over the
//            precomputed distances, so match sets are bit-exact including the stateful gates.
#ifndef ORBSLAM2_B200_ORBMATCHER_H
#define ORBSLAM2_B200_ORBMATCHER_H

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>) && !defined(ORB_B200_NO_OPENCV)
#include <opencv2/core/core.hpp>
#else
#include "cv_compat.h"
#endif
#else
#include <opencv2/core/core.hpp>
#endif

#include "../orb_b200.h"

namespace ORB_SLAM2 {

namespace b200_detail {

// Float arithmetic with every intermediate rounded to float (no fused multiply-add whatever -ffp-contract / -march the
// including translation unit uses): the canonical evaluation of the reference's float expressions.
inline float mulf(float a, float b) { volatile float r = a * b; return r; }
inline float addf(float a, float b) { volatile float r = a + b; return r; }
inline float subf(float a, float b) { volatile float r = a - b; return r; }
inline float divf(float a, float b) { volatile float r = a / b; return r; }

struct Vec3 { float v[3]; float operator[](int i) const { return v[i]; } };
struct Rot3 { float m[9]; };

template <class M> inline Vec3 column3(const M& m, int c = 0) { Vec3 r = {{m.template at<float>(0, c), m.template at<float>(1, c), m.template at<float>(2, c)}}; return r; }
template <class M> inline Rot3 rotation3(const M& m) {
    Rot3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[3 * i + j] = m.template at<float>(i, j);
    return r;
}
// cv::Mat R*p + t for 3x3 * 3x1 CV_32F: OpenCV's small-matrix product sums left to right in float, then the matrix add
inline Vec3 transform(const Rot3& R, const Vec3& p, const Vec3& t) {
    Vec3 r;
    for (int i = 0; i < 3; ++i)
        r.v[i] = addf(addf(addf(mulf(R.m[3 * i], p[0]), mulf(R.m[3 * i + 1], p[1])), mulf(R.m[3 * i + 2], p[2])), t[i]);
    return r;
}
// cv::Mat -R.t()*t: a product with a transposed operand goes through the general gemm, which accumulates floats in double
inline Vec3 minusRtT(const Rot3& R, const Vec3& t) {
    Vec3 r;
    for (int i = 0; i < 3; ++i) {
        volatile double s = (double)R.m[i] * (double)t[0];
        s = s + (double)R.m[3 + i] * (double)t[1];
        s = s + (double)R.m[6 + i] * (double)t[2];
        r.v[i] = (float)(-1.0 * s);
    }
    return r;
}
inline Vec3 minus3(const Vec3& a, const Vec3& b) { Vec3 r = {{subf(a[0], b[0]), subf(a[1], b[1]), subf(a[2], b[2])}}; return r; }
// cv::Mat::dot / cv::norm of 3x1 CV_32F: products and sums in double, in order
inline double dot3(const Vec3& a, const Vec3& b) {
    volatile double s = (double)a[0] * (double)b[0];
    s = s + (double)a[1] * (double)b[1];
    s = s + (double)a[2] * (double)b[2];
    return s;
}
inline double norm3(const Vec3& a) { return std::sqrt(dot3(a, a)); }

// The (query, candidate) pairs a search gathers on the host; one orbm_list_distances launch computes every distance.
class GatedPairs {
public:
    explicit GatedPairs(int device) : device_(device), offsets_(1, 0) {}
    int open(const cv::Mat& desc) {  // a new query row (32 bytes)
        const unsigned char* p = desc.ptr();
        queries_.insert(queries_.end(), p, p + 32);
        return (int)offsets_.size() - 1;
    }
    template <class Indices> void add(const Indices& v) { for (size_t k = 0; k < v.size(); ++k) cands_.push_back((int32_t)v[k]); }
    void close() { offsets_.push_back((int32_t)cands_.size()); }
    void run(const cv::Mat& target) {
        dist_.resize(cands_.size());
        if (cands_.empty()) return;
        std::vector<uint8_t> B((size_t)target.rows * 32);
        for (int i = 0; i < target.rows; ++i) std::memcpy(&B[(size_t)i * 32], target.ptr(i), 32);
        const int rc = orbm_list_distances(device_, queries_.data(), (int)offsets_.size() - 1, B.data(), target.rows, offsets_.data(),
                                           cands_.data(), dist_.data());
        if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error());
    }
    int begin(int q) const { return offsets_[q]; }
    int end(int q) const { return offsets_[q + 1]; }
    size_t cand(int k) const { return (size_t)cands_[k]; }
    int dist(int k) const { return dist_[k]; }
private:
    int device_;
    std::vector<uint8_t> queries_;
    std::vector<int32_t> offsets_, cands_;
    std::vector<int16_t> dist_;
};

// Running nearest / second-nearest of one candidate scan. Every search of the reference keeps them with strict '<' updates
// in visiting order, so the nearest is the FIRST minimum and a tie with it becomes the second; an
// optional tag (the candidate's pyramid level) travels with each of the two.
struct NearestTwo {
    int d1, d2, arg, tag1, tag2;
    explicit NearestTwo(int start) : d1(start), d2(start), arg(-1), tag1(-1), tag2(-1) {}
    void visit(int d, size_t index, int tag = -1) {
        if (d < d1) { d2 = d1; tag2 = tag1; d1 = d; tag1 = tag; arg = (int)index; }
        else if (d < d2) { d2 = d; tag2 = tag; }
    }
    bool passes_ratio(float ratio) const { return static_cast<float>(d1) < ratio * static_cast<float>(d2); }
};
struct Nearest {
    int d, arg;
    explicit Nearest(int start) : d(start), arg(-1) {}
    void visit(int dist, size_t index) { if (dist < d) { d = dist; arg = (int)index; } }
};

// The orientation-consistency filter shared by seven searches: every accepted match votes with the angle difference of its
// two keypoints into one of 30 bins; afterwards only the matches of the three fullest bins survive (second / third bin only
// if they hold at least a tenth of the fullest). bin = round(rot / 30) with rot in [0, 360): the reference's arithmetic,
// which only ever fills bins 0..12 (src/ORBmatcher.cc:240-250, 1603-1644) - kept as it is.
class OrientationVote {
public:
    enum { kBins = 30 };
    static int bin_of(float angle1, float angle2) {
        float rot = angle1 - angle2;
        if (rot < 0.0) rot += 360.0f;
        const int bin = (int)std::round(rot * (1.0f / kBins));
        return bin == kBins ? 0 : bin;
    }
    void cast(float angle1, float angle2, int token) { votes_[bin_of(angle1, angle2)].push_back(token); }
    // the three dominant bins (-1 = none), exactly ComputeThreeMaxima
    void dominant(int& first, int& second, int& third) const {
        int m1 = 0, m2 = 0, m3 = 0;
        first = second = third = -1;
        for (int b = 0; b < kBins; ++b) {
            const int n = (int)votes_[b].size();
            if (n > m1) { m3 = m2; m2 = m1; m1 = n; third = second; second = first; first = b; }
            else if (n > m2) { m3 = m2; m2 = n; third = second; second = b; }
            else if (n > m3) { m3 = n; third = b; }
        }
        if (m2 < 0.1f * (float)m1) { second = -1; third = -1; }
        else if (m3 < 0.1f * (float)m1) third = -1;
    }
    // calls drop(token) for every vote outside the dominant bins; drop returns how many matches that removed (0 or 1)
    template <class Drop> int prune(Drop drop) const {
        int a, b, c, removed = 0;
        dominant(a, b, c);
        for (int i = 0; i < kBins; ++i) {
            if (i == a || i == b || i == c) continue;
            for (size_t j = 0; j < votes_[i].size(); ++j) removed += drop(votes_[i][j]);
        }
        return removed;
    }
private:
    std::vector<int> votes_[kBins];
};

// Does the frame type carry the undistorted image bounds as (static) members like the reference's Frame (include/Frame.h:190-193)?
// Then its grid can be rebuilt on the device and the window gates of a search run there (orbm_window_lists).
template <class F, class = void> struct has_image_bounds : std::false_type {};
template <class F>
struct has_image_bounds<F, decltype((void)(F::mnMinX + F::mnMaxX + F::mnMinY + F::mnMaxY))> : std::true_type {};

}  // namespace b200_detail

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), device_(device) {}

    // Computes the Hamming distance between two ORB descriptors (src/ORBmatcher.cc:1649-1665); host inline
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        const uint32_t* pa = a.ptr<uint32_t>();
        const uint32_t* pb = b.ptr<uint32_t>();
        int dist = 0;
        for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
        return dist;
    }

    // Matching for the map initialization (src/ORBmatcher.cc:407-522). FrameT needs mvKeysUn, mDescriptors and
    // GetFeaturesInArea(x, y, r, minLevel, maxLevel) - the reference's Frame has exactly these.
    template <class FrameT>
    int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10) {
        int nmatches = 0;
        vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);
        b200_detail::OrientationVote vote;
        std::vector<int> claimed_at(F2.mvKeysUn.size(), INT_MAX);   // distance at which a keypoint of F2 is currently matched
        std::vector<int> owner(F2.mvKeysUn.size(), -1);             // ... and by which keypoint of F1

        // 1. + 2. gate and distances. Candidate lists in the reference's iteration order, CSR.
        const size_t n1 = F1.mvKeysUn.size();
        std::vector<int32_t> offsets(n1 + 1, 0), cands;
        std::vector<int16_t> dist;
        if (!window_lists_on_device(F1, F2, vbPrevMatched, windowSize, offsets, cands, dist)) {
            // host gate (frame types without image bounds): the caller's own GetFeaturesInArea, then one launch for all distances
            for (size_t i1 = 0; i1 < n1; i1++) {
                if (F1.mvKeysUn[i1].octave <= 0) {
                    std::vector<size_t> v = F2.GetFeaturesInArea(vbPrevMatched[i1].x, vbPrevMatched[i1].y, windowSize, 0, 0);
                    for (size_t k = 0; k < v.size(); ++k) cands.push_back((int32_t)v[k]);
                }
                offsets[i1 + 1] = (int32_t)cands.size();
            }
            dist.resize(cands.size());
            if (!cands.empty()) {
                const std::vector<uint8_t> A = pack(F1.mDescriptors), B = pack(F2.mDescriptors);
                const int rc = orbm_list_distances(device_, A.data(), (int)n1, B.data(), (int)F2.mvKeysUn.size(), offsets.data(), cands.data(), dist.data());
                if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error());
            }
        }
        // 3. host: the reference's ordered resolve (src/ORBmatcher.cc:434-488)
        for (size_t i1 = 0; i1 < n1; i1++) {
            if (offsets[i1] == offsets[i1 + 1]) continue;
            b200_detail::NearestTwo nn(INT_MAX);
            for (int k = offsets[i1]; k < offsets[i1 + 1]; ++k)
                if (claimed_at[cands[k]] > dist[k]) nn.visit(dist[k], (size_t)cands[k]);   // a closer earlier claim hides the candidate
            if (nn.d1 > TH_LOW || !nn.passes_ratio(mfNNratio)) continue;
            if (owner[nn.arg] >= 0) { vnMatches12[owner[nn.arg]] = -1; nmatches--; }      // take the keypoint over
            vnMatches12[i1] = nn.arg;
            owner[nn.arg] = (int)i1;
            claimed_at[nn.arg] = nn.d1;
            nmatches++;
            if (mbCheckOrientation) vote.cast(F1.mvKeysUn[i1].angle, F2.mvKeysUn[nn.arg].angle, (int)i1);
        }
        if (mbCheckOrientation)
            nmatches -= vote.prune([&](int i1) { if (vnMatches12[i1] < 0) return 0; vnMatches12[i1] = -1; return 1; });
        for (size_t i1 = 0; i1 < vnMatches12.size(); i1++)
            if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;
        return nmatches;
    }

    // Search matches between MapPoints in two keyframes, constrained to features of the same vocabulary node
    // (src/ORBmatcher.cc:524-657) - the matcher of MapFusion::ComputeSim3 / CovisibilityDiscovery
    // (src/MapFusion.cc:275, 849) and of LoopClosing (src/LoopClosing.cc:288). KeyFrameT needs mvKeysUn, mFeatVec
    // (DBoW2::FeatureVector or any ordered map node -> vector<unsigned>), GetMapPointMatches() and mDescriptors;
    // MapPointT needs isBad(). Same split: host merge-walk of the two feature vectors, one device launch for all
    // distances inside equal nodes, the reference's ordered resolve with its vbMatched2 state on the host.
    template <class KeyFrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12) {
        const auto& vKeysUn1 = pKF1->mvKeysUn;
        const auto& vFeatVec1 = pKF1->mFeatVec;
        const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches();
        const cv::Mat& Descriptors1 = pKF1->mDescriptors;
        const auto& vKeysUn2 = pKF2->mvKeysUn;
        const auto& vFeatVec2 = pKF2->mFeatVec;
        const std::vector<MapPointT*> vpMapPoints2 = pKF2->GetMapPointMatches();
        const cv::Mat& Descriptors2 = pKF2->mDescriptors;

        vpMatches12 = std::vector<MapPointT*>(vpMapPoints1.size(), static_cast<MapPointT*>(NULL));
        std::vector<bool> taken2(vpMapPoints2.size(), false);
        b200_detail::OrientationVote vote;
        int nmatches = 0;

        // 1. host: merge walk (552-634) -> CSR of (idx1 occurrence) x (features of the same node in KF2)
        std::vector<int32_t> qrow, offsets(1, 0), cands;
        auto f1it = vFeatVec1.begin(), f1end = vFeatVec1.end();
        auto f2it = vFeatVec2.begin(), f2end = vFeatVec2.end();
        while (f1it != f1end && f2it != f2end) {
            if (f1it->first == f2it->first) {
                for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                    qrow.push_back((int32_t)f1it->second[i1]);
                    for (size_t i2 = 0; i2 < f2it->second.size(); i2++) cands.push_back((int32_t)f2it->second[i2]);
                    offsets.push_back((int32_t)cands.size());
                }
                f1it++; f2it++;
            } else if (f1it->first < f2it->first) f1it = vFeatVec1.lower_bound(f2it->first);
            else f2it = vFeatVec2.lower_bound(f1it->first);
        }
        // 2. device: all distances
        std::vector<int16_t> dist(cands.size());
        if (!cands.empty()) {
            std::vector<uint8_t> A(qrow.size() * 32);
            for (size_t q = 0; q < qrow.size(); ++q) std::memcpy(&A[q * 32], Descriptors1.ptr(qrow[q]), 32);
            const std::vector<uint8_t> B = pack(Descriptors2);
            const int rc = orbm_list_distances(device_, A.data(), (int)qrow.size(), B.data(), Descriptors2.rows, offsets.data(), cands.data(), dist.data());
            if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error());
        }
        // 3. host: ordered resolve (556-620)
        for (size_t q = 0; q < qrow.size(); ++q) {
            const size_t idx1 = (size_t)qrow[q];
            MapPointT* pMP1 = vpMapPoints1[idx1];
            if (!pMP1) continue;
            if (pMP1->isBad()) continue;
            b200_detail::NearestTwo nn(256);
            for (int k = offsets[q]; k < offsets[q + 1]; ++k) {
                const size_t idx2 = (size_t)cands[k];
                MapPointT* pMP2 = vpMapPoints2[idx2];
                if (taken2[idx2] || !pMP2 || pMP2->isBad()) continue;
                nn.visit(dist[k], idx2);
            }
            if (nn.d1 >= TH_LOW || !nn.passes_ratio(mfNNratio)) continue;   // strictly below TH_LOW here (600)
            vpMatches12[idx1] = vpMapPoints2[nn.arg];
            taken2[nn.arg] = true;
            if (mbCheckOrientation) vote.cast(vKeysUn1[idx1].angle, vKeysUn2[nn.arg].angle, (int)idx1);
            nmatches++;
        }
        if (mbCheckOrientation) nmatches -= vote.prune([&](int idx1) { vpMatches12[idx1] = static_cast<MapPointT*>(NULL); return 1; });
        return nmatches;
    }

    // ---- the projection / BoW guided searches (SURVEY.md section 8 rows a-10 ... a-15) --------------------------------
    // All of them: (1) host: the reference's outer loop with every gate that does not depend on matches made earlier in the
    // same call, collecting the GetFeaturesInArea / BoW-node candidates in the reference's order; (2) device: one
    // orbm_list_distances launch; (3) host: the reference's ordered selection replayed over the precomputed distances on
    // the caller's live objects (so mvpMapPoints / vpMatched state behaves exactly as in the reference).

    // Search matches between Frame keypoints and projected MapPoints (src/ORBmatcher.cc:45-131); used to track the local
    // map (Tracking::SearchLocalPoints). FrameT: mpSystem, mvScaleFactors, GetFeaturesInArea, mvpMapPoints, mvuRight,
    // mDescriptors, mvKeysUn. MapPointT: mbTrackInView / mnTrackScaleLevel / mTrackViewCos / mTrackProjX / mTrackProjY /
    // mTrackProjXR (per System*), isBad(), GetDescriptor(), Observations().
    template <class FrameT, class MapPointT>
    int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float th = 3) {
        struct Query { MapPointT* mp; int level; float radius; int q; };
        std::vector<Query> queries;
        b200_detail::GatedPairs pairs(device_);
        const bool bFactor = th != 1.0;
        auto pSystem = F.mpSystem;
        for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
            MapPointT* pMP = vpMapPoints[iMP];
            if (!pMP->mbTrackInView[pSystem] || pMP->isBad()) continue;
            const int nPredictedLevel = pMP->mnTrackScaleLevel[pSystem];
            float r = RadiusByViewingCos(pMP->mTrackViewCos[pSystem]);  // window size depends on the viewing direction
            if (bFactor) r *= th;
            const float radius = r * F.mvScaleFactors[nPredictedLevel];
            const std::vector<size_t> vIndices =
                F.GetFeaturesInArea(pMP->mTrackProjX[pSystem], pMP->mTrackProjY[pSystem], radius, nPredictedLevel - 1, nPredictedLevel);
            if (vIndices.empty()) continue;
            const Query q = {pMP, nPredictedLevel, radius, pairs.open(pMP->GetDescriptor())};
            pairs.add(vIndices);
            pairs.close();
            queries.push_back(q);
        }
        pairs.run(F.mDescriptors);
        int nmatches = 0;
        for (size_t n = 0; n < queries.size(); ++n) {
            const Query& q = queries[n];
            b200_detail::NearestTwo nn(256);
            for (int k = pairs.begin(q.q); k < pairs.end(q.q); ++k) {
                const size_t idx = pairs.cand(k);
                if (F.mvpMapPoints[idx] && F.mvpMapPoints[idx]->Observations() > 0) continue;
                if (F.mvuRight[idx] > 0) {
                    const float er = fabs(q.mp->mTrackProjXR[pSystem] - F.mvuRight[idx]);
                    if (er > q.radius) continue;
                }
                nn.visit(pairs.dist(k), idx, F.mvKeysUn[idx].octave);
            }
            if (nn.d1 > TH_HIGH) continue;
            // the ratio to the second match counts only when both sit on the same pyramid level (119-122)
            if (nn.tag1 == nn.tag2 && nn.d1 > mfNNratio * nn.d2) continue;
            F.mvpMapPoints[nn.arg] = q.mp;
            nmatches++;
        }
        return nmatches;
    }

    // Project MapPoints tracked in the last frame into the current frame and search matches (src/ORBmatcher.cc:1330-1472);
    // used to track from the previous frame (Tracking::TrackWithMotionModel).
    template <class FrameT>
    int SearchByProjection(FrameT& CurrentFrame, const FrameT& LastFrame, const float th, const bool bMono) {
        using namespace b200_detail;
        typedef typename std::remove_cv<typename std::remove_reference<decltype(CurrentFrame.mvpMapPoints[0])>::type>::type MapPointPtr;
        const Rot3 Rcw = rotation3(CurrentFrame.mTcw);
        const Vec3 tcw = column3(CurrentFrame.mTcw, 3);
        const Vec3 twc = minusRtT(Rcw, tcw);
        const Vec3 tlc = transform(rotation3(LastFrame.mTcw), twc, column3(LastFrame.mTcw, 3));
        const bool bForward = tlc[2] > CurrentFrame.mb && !bMono;
        const bool bBackward = -tlc[2] > CurrentFrame.mb && !bMono;

        struct Query { int i; MapPointPtr mp; float ur, radius; int q; };
        std::vector<Query> queries;
        GatedPairs pairs(device_);
        for (int i = 0; i < LastFrame.N; i++) {
            MapPointPtr pMP = LastFrame.mvpMapPoints[i];
            if (!pMP || LastFrame.mvbOutlier[i]) continue;
            const Vec3 x3Dc = transform(Rcw, column3(pMP->GetWorldPos()), tcw);
            const float invzc = 1.0 / x3Dc[2];
            if (invzc < 0) continue;
            const float u = addf(mulf(mulf(CurrentFrame.fx, x3Dc[0]), invzc), CurrentFrame.cx);
            const float v = addf(mulf(mulf(CurrentFrame.fy, x3Dc[1]), invzc), CurrentFrame.cy);
            if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
            if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
            const int nLastOctave = LastFrame.mvKeys[i].octave;
            const float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];  // window size depends on the scale
            std::vector<size_t> vIndices2;
            if (bForward) vIndices2 = CurrentFrame.GetFeaturesInArea(u, v, radius, nLastOctave);
            else if (bBackward) vIndices2 = CurrentFrame.GetFeaturesInArea(u, v, radius, 0, nLastOctave);
            else vIndices2 = CurrentFrame.GetFeaturesInArea(u, v, radius, nLastOctave - 1, nLastOctave + 1);
            if (vIndices2.empty()) continue;
            const Query q = {i, pMP, subf(u, mulf(CurrentFrame.mbf, invzc)), radius, pairs.open(pMP->GetDescriptor())};
            pairs.add(vIndices2);
            pairs.close();
            queries.push_back(q);
        }
        pairs.run(CurrentFrame.mDescriptors);

        int nmatches = 0;
        OrientationVote vote;
        for (size_t n = 0; n < queries.size(); ++n) {
            const Query& q = queries[n];
            Nearest nn(256);
            for (int k = pairs.begin(q.q); k < pairs.end(q.q); ++k) {
                const size_t i2 = pairs.cand(k);
                if (CurrentFrame.mvpMapPoints[i2] && CurrentFrame.mvpMapPoints[i2]->Observations() > 0) continue;
                if (CurrentFrame.mvuRight[i2] > 0) {
                    const float er = fabs(q.ur - CurrentFrame.mvuRight[i2]);
                    if (er > q.radius) continue;
                }
                nn.visit(pairs.dist(k), i2);
            }
            if (nn.d > TH_HIGH) continue;
            CurrentFrame.mvpMapPoints[nn.arg] = q.mp;
            nmatches++;
            if (mbCheckOrientation) vote.cast(LastFrame.mvKeysUn[q.i].angle, CurrentFrame.mvKeysUn[nn.arg].angle, nn.arg);
        }
        if (mbCheckOrientation) nmatches -= vote.prune([&](int i2) { CurrentFrame.mvpMapPoints[i2] = static_cast<MapPointPtr>(NULL); return 1; });
        return nmatches;
    }

    // Project MapPoints seen in a KeyFrame into the Frame and search matches (src/ORBmatcher.cc:1474-1601); relocalisation.
    template <class FrameT, class KeyFrameT, class MapPointT>
    int SearchByProjection(FrameT& CurrentFrame, KeyFrameT* pKF, const std::set<MapPointT*>& sAlreadyFound, const float th, const int ORBdist) {
        using namespace b200_detail;
        const Rot3 Rcw = rotation3(CurrentFrame.mTcw);
        const Vec3 tcw = column3(CurrentFrame.mTcw, 3);
        const Vec3 Ow = minusRtT(Rcw, tcw);
        const std::vector<MapPointT*> vpMPs = pKF->GetMapPointMatches();

        struct Query { size_t i; MapPointT* mp; int q; };
        std::vector<Query> queries;
        GatedPairs pairs(device_);
        for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
            MapPointT* pMP = vpMPs[i];
            if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
            const Vec3 x3Dw = column3(pMP->GetWorldPos());
            const Vec3 x3Dc = transform(Rcw, x3Dw, tcw);
            const float invzc = 1.0 / x3Dc[2];
            const float u = addf(mulf(mulf(CurrentFrame.fx, x3Dc[0]), invzc), CurrentFrame.cx);
            const float v = addf(mulf(mulf(CurrentFrame.fy, x3Dc[1]), invzc), CurrentFrame.cy);
            if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
            if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
            float dist3D = norm3(minus3(x3Dw, Ow));  // depth must be inside the scale pyramid of the image
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            const int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
            const float radius = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
            const std::vector<size_t> vIndices2 = CurrentFrame.GetFeaturesInArea(u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1);
            if (vIndices2.empty()) continue;
            const Query q = {i, pMP, pairs.open(pMP->GetDescriptor())};
            pairs.add(vIndices2);
            pairs.close();
            queries.push_back(q);
        }
        pairs.run(CurrentFrame.mDescriptors);

        int nmatches = 0;
        OrientationVote vote;
        for (size_t n = 0; n < queries.size(); ++n) {
            const Query& q = queries[n];
            Nearest nn(256);
            for (int k = pairs.begin(q.q); k < pairs.end(q.q); ++k)
                if (!CurrentFrame.mvpMapPoints[pairs.cand(k)]) nn.visit(pairs.dist(k), pairs.cand(k));
            if (nn.d > ORBdist) continue;
            CurrentFrame.mvpMapPoints[nn.arg] = q.mp;
            nmatches++;
            if (mbCheckOrientation) vote.cast(pKF->mvKeysUn[q.i].angle, CurrentFrame.mvKeysUn[nn.arg].angle, nn.arg);
        }
        if (mbCheckOrientation) nmatches -= vote.prune([&](int i2) { CurrentFrame.mvpMapPoints[i2] = static_cast<MapPointT*>(NULL); return 1; });
        return nmatches;
    }

    // Project MapPoints using a Similarity Transformation and search matches (src/ORBmatcher.cc:292-405); loop detection.
    template <class KeyFrameT, class MapPointT>
    int SearchByProjection(KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched, int th) {
        using namespace b200_detail;
        const Sim3Camera cam = DecomposeSim3(Scw);
        std::set<MapPointT*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
        spAlreadyFound.erase(static_cast<MapPointT*>(NULL));

        struct Query { MapPointT* mp; int q; };
        std::vector<Query> queries;
        GatedPairs pairs(device_);
        for (int iMP = 0, iendMP = (int)vpPoints.size(); iMP < iendMP; iMP++) {
            MapPointT* pMP = vpPoints[iMP];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            const Vec3 p3Dw = column3(pMP->GetWorldPos());
            const Vec3 p3Dc = transform(cam.R, p3Dw, cam.t);
            if (p3Dc[2] < 0.0) continue;
            const float invz = divf(1.0f, p3Dc[2]);
            const float u = addf(mulf(pKF->fx, mulf(p3Dc[0], invz)), pKF->cx);
            const float v = addf(mulf(pKF->fy, mulf(p3Dc[1], invz)), pKF->cy);
            if (!pKF->IsInImage(u, v)) continue;
            const Vec3 PO = minus3(p3Dw, cam.Ow);
            const float dist = norm3(PO);
            if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) continue;
            if (dot3(PO, column3(pMP->GetNormal())) < 0.5 * dist) continue;  // viewing angle must be less than 60 deg
            const int nPredictedLevel = pMP->PredictScale(dist, pKF);
            const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
            const std::vector<size_t> vIndices = pKF->GetFeaturesInArea(u, v, radius);
            if (vIndices.empty()) continue;
            const Query q = {pMP, pairs.open(pMP->GetDescriptor())};
            pairs.add(LevelWindow(vIndices, pKF->mvKeysUn, nPredictedLevel));
            pairs.close();
            queries.push_back(q);
        }
        pairs.run(pKF->mDescriptors);

        int nmatches = 0;
        for (size_t n = 0; n < queries.size(); ++n) {
            const Query& q = queries[n];
            Nearest nn(256);
            for (int k = pairs.begin(q.q); k < pairs.end(q.q); ++k)
                if (!vpMatched[pairs.cand(k)]) nn.visit(pairs.dist(k), pairs.cand(k));
            if (nn.d <= TH_LOW) { vpMatched[nn.arg] = q.mp; nmatches++; }
        }
        return nmatches;
    }

    // Search matches between MapPoints in a KeyFrame and ORB in a Frame, constrained to features of the same vocabulary
    // node (src/ORBmatcher.cc:161-290); relocalisation and TrackReferenceKeyFrame.
    template <class KeyFrameT, class FrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches) {
        const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches = std::vector<MapPointT*>(F.N, static_cast<MapPointT*>(NULL));
        struct Query { unsigned idxKF; MapPointT* mp; int q; };
        std::vector<Query> queries;
        b200_detail::GatedPairs pairs(device_);
        auto KFit = pKF->mFeatVec.begin(), KFend = pKF->mFeatVec.end();
        auto Fit = F.mFeatVec.begin(), Fend = F.mFeatVec.end();
        while (KFit != KFend && Fit != Fend) {
            if (KFit->first == Fit->first) {
                for (size_t iKF = 0; iKF < KFit->second.size(); iKF++) {
                    const unsigned realIdxKF = KFit->second[iKF];
                    MapPointT* pMP = vpMapPointsKF[realIdxKF];
                    if (!pMP || pMP->isBad()) continue;
                    const Query q = {realIdxKF, pMP, pairs.open(pKF->mDescriptors.row((int)realIdxKF))};
                    pairs.add(Fit->second);
                    pairs.close();
                    queries.push_back(q);
                }
                KFit++; Fit++;
            } else if (KFit->first < Fit->first) KFit = pKF->mFeatVec.lower_bound(Fit->first);
            else Fit = F.mFeatVec.lower_bound(KFit->first);
        }
        pairs.run(F.mDescriptors);

        int nmatches = 0;
        b200_detail::OrientationVote vote;
        for (size_t n = 0; n < queries.size(); ++n) {
            const Query& q = queries[n];
            b200_detail::NearestTwo nn(256);
            for (int k = pairs.begin(q.q); k < pairs.end(q.q); ++k)
                if (!vpMapPointMatches[pairs.cand(k)]) nn.visit(pairs.dist(k), pairs.cand(k));
            if (nn.d1 > TH_LOW || !nn.passes_ratio(mfNNratio)) continue;
            vpMapPointMatches[nn.arg] = q.mp;
            if (mbCheckOrientation) vote.cast(pKF->mvKeysUn[q.idxKF].angle, F.mvKeys[nn.arg].angle, nn.arg);   // F.mvKeys, not Un (235)
            nmatches++;
        }
        if (mbCheckOrientation) nmatches -= vote.prune([&](int iF) { vpMapPointMatches[iF] = static_cast<MapPointT*>(NULL); return 1; });
        return nmatches;
    }

    // Matching to triangulate new MapPoints, with the epipolar constraint (src/ORBmatcher.cc:659-825); LocalMapping.
    template <class KeyFrameT>
    int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo) {
        using namespace b200_detail;
        // epipole in the second image
        const Vec3 C2 = transform(rotation3(pKF2->GetRotation()), column3(pKF1->GetCameraCenter()), column3(pKF2->GetTranslation()));
        const float invz = divf(1.0f, C2[2]);
        const float ex = addf(mulf(mulf(pKF2->fx, C2[0]), invz), pKF2->cx);
        const float ey = addf(mulf(mulf(pKF2->fy, C2[1]), invz), pKF2->cy);

        struct Query { size_t idx1; bool stereo1; int q; };
        std::vector<Query> queries;
        GatedPairs pairs(device_);
        auto f1it = pKF1->mFeatVec.begin(), f1end = pKF1->mFeatVec.end();
        auto f2it = pKF2->mFeatVec.begin(), f2end = pKF2->mFeatVec.end();
        std::vector<size_t> open2;
        while (f1it != f1end && f2it != f2end) {
            if (f1it->first == f2it->first) {
                open2.clear();  // features of this node in KF2 that have no MapPoint yet (and are stereo if required)
                for (size_t i2 = 0; i2 < f2it->second.size(); i2++) {
                    const size_t idx2 = f2it->second[i2];
                    if (pKF2->GetMapPoint(idx2)) continue;
                    if (bOnlyStereo && !(pKF2->mvuRight[idx2] >= 0)) continue;
                    open2.push_back(idx2);
                }
                for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                    const size_t idx1 = f1it->second[i1];
                    if (pKF1->GetMapPoint(idx1)) continue;
                    const bool bStereo1 = pKF1->mvuRight[idx1] >= 0;
                    if (bOnlyStereo && !bStereo1) continue;
                    const Query q = {idx1, bStereo1, pairs.open(pKF1->mDescriptors.row((int)idx1))};
                    pairs.add(open2);
                    pairs.close();
                    queries.push_back(q);
                }
                f1it++; f2it++;
            } else if (f1it->first < f2it->first) f1it = pKF1->mFeatVec.lower_bound(f2it->first);
            else f2it = pKF2->mFeatVec.lower_bound(f1it->first);
        }
        pairs.run(pKF2->mDescriptors);

        int nmatches = 0;
        std::vector<int> partner(pKF1->N, -1);
        OrientationVote vote;
        for (size_t n = 0; n < queries.size(); ++n) {
            const Query& q = queries[n];
            const cv::KeyPoint& kp1 = pKF1->mvKeysUn[q.idx1];
            // not the strict rule here: a candidate as close as the current choice replaces it (`dist > bestDist` skips, 740),
            // provided it passes the geometric gates - so the LAST minimum among the admissible candidates wins
            int limit = TH_LOW, chosen = -1;
            for (int k = pairs.begin(q.q); k < pairs.end(q.q); ++k) {
                const size_t idx2 = pairs.cand(k);
                const int dist = pairs.dist(k);
                if (dist > limit) continue;
                const cv::KeyPoint& kp2 = pKF2->mvKeysUn[idx2];
                if (!q.stereo1 && !(pKF2->mvuRight[idx2] >= 0)) {  // too close to the epipole: skip
                    const float distex = subf(ex, kp2.pt.x), distey = subf(ey, kp2.pt.y);
                    if (addf(mulf(distex, distex), mulf(distey, distey)) < 100 * pKF2->mvScaleFactors[kp2.octave]) continue;
                }
                if (CheckDistEpipolarLine(kp1, kp2, F12, pKF2)) { chosen = (int)idx2; limit = dist; }
            }
            if (chosen < 0) continue;
            partner[q.idx1] = chosen;
            nmatches++;
            if (mbCheckOrientation) vote.cast(kp1.angle, pKF2->mvKeysUn[chosen].angle, (int)q.idx1);
        }
        if (mbCheckOrientation) nmatches -= vote.prune([&](int idx1) { partner[idx1] = -1; return 1; });
        vMatchedPairs.clear();
        vMatchedPairs.reserve(nmatches);
        for (size_t i = 0; i < partner.size(); i++)
            if (partner[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)partner[i]));
        return nmatches;
    }

    // Project MapPoints into a KeyFrame and search for duplicated MapPoints (src/ORBmatcher.cc:827-977); LocalMapping.
    template <class KeyFrameT, class MapPointT>
    int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, const float th = 3.0) {
        using namespace b200_detail;
        const Rot3 Rcw = rotation3(pKF->GetRotation());
        const Vec3 tcw = column3(pKF->GetTranslation());
        const Vec3 Ow = column3(pKF->GetCameraCenter());
        struct Query { MapPointT* mp; int q; };
        std::vector<Query> queries;
        GatedPairs pairs(device_);
        std::vector<size_t> kept;
        for (int i = 0, nMPs = (int)vpMapPoints.size(); i < nMPs; i++) {
            MapPointT* pMP = vpMapPoints[i];
            if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
            const Vec3 p3Dw = column3(pMP->GetWorldPos());
            const Vec3 p3Dc = transform(Rcw, p3Dw, tcw);
            if (p3Dc[2] < 0.0f) continue;
            const float invz = divf(1.0f, p3Dc[2]);
            const float u = addf(mulf(pKF->fx, mulf(p3Dc[0], invz)), pKF->cx);
            const float v = addf(mulf(pKF->fy, mulf(p3Dc[1], invz)), pKF->cy);
            if (!pKF->IsInImage(u, v)) continue;
            const float ur = subf(u, mulf(pKF->mbf, invz));
            const Vec3 PO = minus3(p3Dw, Ow);
            const float dist3D = norm3(PO);
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            if (dot3(PO, column3(pMP->GetNormal())) < 0.5 * dist3D) continue;
            const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
            const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
            const std::vector<size_t> vIndices = pKF->GetFeaturesInArea(u, v, radius);
            if (vIndices.empty()) continue;
            kept.clear();  // level window and reprojection-error gate (stateless): 902-936
            for (size_t n = 0; n < vIndices.size(); ++n) {
                const size_t idx = vIndices[n];
                const cv::KeyPoint& kp = pKF->mvKeysUn[idx];
                const int kpLevel = kp.octave;
                if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
                const float ex = subf(u, kp.pt.x), ey = subf(v, kp.pt.y);
                float e2 = addf(mulf(ex, ex), mulf(ey, ey));
                if (pKF->mvuRight[idx] >= 0) {
                    const float er = subf(ur, pKF->mvuRight[idx]);
                    e2 = addf(e2, mulf(er, er));
                    if (mulf(e2, pKF->mvInvLevelSigma2[kpLevel]) > 7.8) continue;
                } else if (mulf(e2, pKF->mvInvLevelSigma2[kpLevel]) > 5.99) continue;
                kept.push_back(idx);
            }
            const Query q = {pMP, pairs.open(pMP->GetDescriptor())};
            pairs.add(kept);
            pairs.close();
            queries.push_back(q);
        }
        pairs.run(pKF->mDescriptors);

        int nFused = 0;
        for (size_t n = 0; n < queries.size(); ++n) {
            MapPointT* pMP = queries[n].mp;
            if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;  // earlier replacements of this call may have changed it
            Nearest nn(256);
            for (int k = pairs.begin(queries[n].q); k < pairs.end(queries[n].q); ++k) nn.visit(pairs.dist(k), pairs.cand(k));
            const int bestIdx = nn.arg;
            if (nn.d <= TH_LOW) {  // replace if the keypoint already has a MapPoint, otherwise add the measurement
                MapPointT* pMPinKF = pKF->GetMapPoint(bestIdx);
                if (pMPinKF) {
                    if (!pMPinKF->isBad()) {
                        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                        else pMPinKF->Replace(pMP);
                    }
                } else {
                    pMP->AddObservation(pKF, bestIdx);
                    pKF->AddMapPoint(pMP, bestIdx);
                }
                nFused++;
            }
        }
        return nFused;
    }

    // Project MapPoints into a KeyFrame using a given Sim3 and search for duplicated MapPoints (src/ORBmatcher.cc:979-1102).
    template <class KeyFrameT, class MapPointT>
    int Fuse(KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints, float th, std::vector<MapPointT*>& vpReplacePoint) {
        using namespace b200_detail;
        const Sim3Camera cam = DecomposeSim3(Scw);
        const std::set<MapPointT*> spAlreadyFound = pKF->GetMapPoints();
        struct Query { int iMP; MapPointT* mp; int q; };
        std::vector<Query> queries;
        GatedPairs pairs(device_);
        for (int iMP = 0, nPoints = (int)vpPoints.size(); iMP < nPoints; iMP++) {
            MapPointT* pMP = vpPoints[iMP];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            const Vec3 p3Dw = column3(pMP->GetWorldPos());
            const Vec3 p3Dc = transform(cam.R, p3Dw, cam.t);
            if (p3Dc[2] < 0.0f) continue;
            const float invz = 1.0 / p3Dc[2];
            const float u = addf(mulf(pKF->fx, mulf(p3Dc[0], invz)), pKF->cx);
            const float v = addf(mulf(pKF->fy, mulf(p3Dc[1], invz)), pKF->cy);
            if (!pKF->IsInImage(u, v)) continue;
            const Vec3 PO = minus3(p3Dw, cam.Ow);
            const float dist3D = norm3(PO);
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            if (dot3(PO, column3(pMP->GetNormal())) < 0.5 * dist3D) continue;
            const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
            const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
            const std::vector<size_t> vIndices = pKF->GetFeaturesInArea(u, v, radius);
            if (vIndices.empty()) continue;
            const Query q = {iMP, pMP, pairs.open(pMP->GetDescriptor())};
            pairs.add(LevelWindow(vIndices, pKF->mvKeysUn, nPredictedLevel));
            pairs.close();
            queries.push_back(q);
        }
        pairs.run(pKF->mDescriptors);

        int nFused = 0;
        for (size_t n = 0; n < queries.size(); ++n) {
            const Query& q = queries[n];
            Nearest nn(INT_MAX);
            for (int k = pairs.begin(q.q); k < pairs.end(q.q); ++k) nn.visit(pairs.dist(k), pairs.cand(k));
            const int bestIdx = nn.arg;
            if (nn.d <= TH_LOW) {
                MapPointT* pMPinKF = pKF->GetMapPoint(bestIdx);
                if (pMPinKF) {
                    if (!pMPinKF->isBad()) vpReplacePoint[q.iMP] = pMPinKF;
                } else {
                    q.mp->AddObservation(pKF, bestIdx);
                    pKF->AddMapPoint(q.mp, bestIdx);
                }
                nFused++;
            }
        }
        return nFused;
    }

    // Search matches between MapPoints seen in KF1 and KF2, transforming by a Sim3 [s12*R12|t12]
    // (src/ORBmatcher.cc:1104-1328); loop closing and MapFusion::ComputeSim3.
    template <class KeyFrameT, class MapPointT>
    int SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float& s12, const cv::Mat& R12,
                     const cv::Mat& t12, const float th) {
        using namespace b200_detail;
        // transformation between the cameras: sR12 = s12*R12, sR21 = (1.0/s12)*R12.t(), t21 = -sR21*t12 (cv::Mat scaling
        // multiplies every element by the float of the double factor)
        Rot3 sR12 = rotation3(R12), sR21;
        const float inv_s = (float)(1.0 / s12);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sR21.m[3 * i + j] = mulf(sR12.m[3 * j + i], inv_s);
        for (int i = 0; i < 9; ++i) sR12.m[i] = mulf(sR12.m[i], s12);
        const Vec3 t12v = column3(t12), zero = {{0.f, 0.f, 0.f}};
        Vec3 t21 = transform(sR21, t12v, zero);
        for (int i = 0; i < 3; ++i) t21.v[i] = -t21.v[i];

        const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
        std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
        for (int i = 0; i < N1; i++) {
            MapPointT* pMP = vpMatches12[i];
            if (!pMP) continue;
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
        const std::vector<int> vnMatch1 = Sim3Direction(pKF1, pKF2, pKF1, vpMapPoints1, vbAlreadyMatched1, sR21, t21, th);  // KF1 -> KF2
        const std::vector<int> vnMatch2 = Sim3Direction(pKF2, pKF1, pKF1, vpMapPoints2, vbAlreadyMatched2, sR12, t12v, th);  // KF2 -> KF1
        int nFound = 0;  // check agreement
        for (int i1 = 0; i1 < N1; i1++) {
            const int idx2 = vnMatch1[i1];
            if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
        }
        return nFound;
    }

    // Brute-force ratio-test matching of two descriptor matrices (the SearchByBoW(KF,KF) inner loop,
    // src/ORBmatcher.cc:566-603, with the vocabulary gate removed): vnMatches12[i] = row of D2 or -1.
    int SearchBruteForce(const cv::Mat& D1, const cv::Mat& D2, std::vector<int>& vnMatches12, int th = TH_LOW) {
        const int n1 = D1.rows, n2 = D2.rows;
        vnMatches12.assign(n1, -1);
        if (n1 == 0 || n2 == 0) return 0;
        const std::vector<uint8_t> A = pack(D1), B = pack(D2);
        std::vector<int32_t> idx(n1), b1(n1), b2(n1);
        const int rc = orbm_knn2(device_, A.data(), n1, B.data(), n2, idx.data(), b1.data(), b2.data());
        if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error());
        int n = 0;
        for (int i = 0; i < n1; ++i)
            if (b1[i] < th && static_cast<float>(b1[i]) < mfNNratio * static_cast<float>(b2[i])) { vnMatches12[i] = idx[i]; ++n; }
        return n;
    }

    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

protected:
    static float RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }  // src/ORBmatcher.cc:133-139

    // the rotation-histogram slot of a match (e.g. src/ORBmatcher.cc:1434-1441): bin = round(rot * 1/HISTO_LENGTH)
    static int RotationBin(float angle1, float angle2) {
        float rot = angle1 - angle2;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * (1.0f / HISTO_LENGTH));
        if (bin == HISTO_LENGTH) bin = 0;
        return bin;
    }

    // candidates at the levels [nPredictedLevel-1, nPredictedLevel], order kept (e.g. src/ORBmatcher.cc:375-378)
    template <class KeyPoints>
    static std::vector<size_t> LevelWindow(const std::vector<size_t>& vIndices, const KeyPoints& vKeysUn, int nPredictedLevel) {
        std::vector<size_t> kept;
        for (size_t n = 0; n < vIndices.size(); ++n) {
            const int kpLevel = vKeysUn[vIndices[n]].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            kept.push_back(vIndices[n]);
        }
        return kept;
    }

    // "Decompose Scw" (src/ORBmatcher.cc:300-305, 987-992): scw = |first row of sR|, Rcw = sRcw/scw, tcw = t/scw, Ow = -Rcw.t()*tcw.
    // cv::Mat / scalar multiplies every element by the float of the double reciprocal.
    struct Sim3Camera { b200_detail::Rot3 R; b200_detail::Vec3 t, Ow; };
    static Sim3Camera DecomposeSim3(const cv::Mat& Scw) {
        using namespace b200_detail;
        Sim3Camera c;
        const Rot3 sR = rotation3(Scw);
        const Vec3 row0 = {{sR.m[0], sR.m[1], sR.m[2]}};
        const float scw = std::sqrt(dot3(row0, row0));
        const float inv = (float)(1.0 / scw);
        for (int i = 0; i < 9; ++i) c.R.m[i] = mulf(sR.m[i], inv);
        const Vec3 t = column3(Scw, 3);
        for (int i = 0; i < 3; ++i) c.t.v[i] = mulf(t[i], inv);
        c.Ow = minusRtT(c.R, c.t);
        return c;
    }

    // Epipolar line in the second image l = x1'F12 = [a b c]; distance test (src/ORBmatcher.cc:142-159)
    template <class KeyFrameT>
    static bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrameT* pKF2) {
        using namespace b200_detail;
        const float a = addf(addf(mulf(kp1.pt.x, F12.at<float>(0, 0)), mulf(kp1.pt.y, F12.at<float>(1, 0))), F12.at<float>(2, 0));
        const float b = addf(addf(mulf(kp1.pt.x, F12.at<float>(0, 1)), mulf(kp1.pt.y, F12.at<float>(1, 1))), F12.at<float>(2, 1));
        const float c = addf(addf(mulf(kp1.pt.x, F12.at<float>(0, 2)), mulf(kp1.pt.y, F12.at<float>(1, 2))), F12.at<float>(2, 2));
        const float num = addf(addf(mulf(a, kp2.pt.x), mulf(b, kp2.pt.y)), c);
        const float den = addf(mulf(a, a), mulf(b, b));
        if (den == 0) return false;
        const float dsqr = divf(mulf(num, num), den);
        return dsqr < 3.84 * pKF2->mvLevelSigma2[kp2.octave];
    }

    // One direction of SearchBySim3 (src/ORBmatcher.cc:1152-1229 / 1232-1309): the MapPoints of pKFa transformed into pKFb
    // by [sRba|tba] after their own camera transform, projected with pKFcal's intrinsics; no state between queries, so the device result is used as is.
    template <class KeyFrameT, class MapPointT>
    std::vector<int> Sim3Direction(KeyFrameT* pKFa, KeyFrameT* pKFb, KeyFrameT* pKFcal, const std::vector<MapPointT*>& vpMapPointsA,
                                   const std::vector<bool>& vbAlreadyMatchedA, const b200_detail::Rot3& sRba, const b200_detail::Vec3& tba,
                                   const float th) {
        using namespace b200_detail;
        const Rot3 Raw = rotation3(pKFa->GetRotation());
        const Vec3 taw = column3(pKFa->GetTranslation());
        const int NA = (int)vpMapPointsA.size();
        std::vector<int> vnMatch(NA, -1), owner;
        GatedPairs pairs(device_);
        for (int i = 0; i < NA; i++) {
            MapPointT* pMP = vpMapPointsA[i];
            if (!pMP || vbAlreadyMatchedA[i] || pMP->isBad()) continue;
            const Vec3 p3Dca = transform(Raw, column3(pMP->GetWorldPos()), taw);
            const Vec3 p3Dcb = transform(sRba, p3Dca, tba);
            if (p3Dcb[2] < 0.0) continue;
            const float invz = 1.0 / p3Dcb[2];
            const float u = addf(mulf(pKFcal->fx, mulf(p3Dcb[0], invz)), pKFcal->cx);  // pKF1's intrinsics in both directions (1107-1110)
            const float v = addf(mulf(pKFcal->fy, mulf(p3Dcb[1], invz)), pKFcal->cy);
            if (!pKFb->IsInImage(u, v)) continue;
            const float dist3D = norm3(p3Dcb);
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            const int nPredictedLevel = pMP->PredictScale(dist3D, pKFb);
            const float radius = th * pKFb->mvScaleFactors[nPredictedLevel];
            const std::vector<size_t> vIndices = pKFb->GetFeaturesInArea(u, v, radius);
            if (vIndices.empty()) continue;
            pairs.open(pMP->GetDescriptor());
            pairs.add(LevelWindow(vIndices, pKFb->mvKeysUn, nPredictedLevel));
            pairs.close();
            owner.push_back(i);
        }
        pairs.run(pKFb->mDescriptors);
        for (size_t q = 0; q < owner.size(); ++q) {
            int bestDist = INT_MAX, bestIdx = -1;
            for (int k = pairs.begin((int)q); k < pairs.end((int)q); ++k)
                if (pairs.dist(k) < bestDist) { bestDist = pairs.dist(k); bestIdx = (int)pairs.cand(k); }
            if (bestDist <= TH_HIGH) vnMatch[owner[q]] = bestIdx;
        }
        return vnMatch;
    }

    // SearchForInitialization's gate on the device, for frame types that expose the image bounds (the reference's Frame):
    // F2's keypoints + descriptors and the level-0 windows of F1 go up in one copy, the device rebuilds F2's grid
    // (AssignFeaturesToGrid, src/Frame.cc:230-245), walks every window in GetFeaturesInArea's order (327-380) and returns
    // candidates + Hamming distances - instead of ~3 us of host grid walking per 100 px window plus a separate distance launch.
    template <class FrameT>
    typename std::enable_if<b200_detail::has_image_bounds<FrameT>::value, bool>::type
    window_lists_on_device(FrameT& F1, FrameT& F2, const std::vector<cv::Point2f>& centres, int windowSize, std::vector<int32_t>& offsets,
                           std::vector<int32_t>& cands, std::vector<int16_t>& dist) {
        const size_t n1 = F1.mvKeysUn.size(), n2 = F2.mvKeysUn.size();
        if (n1 == 0 || n2 == 0) return true;   // offsets are all zero already
        std::vector<orbx_keypoint> kp2(n2);
        for (size_t i = 0; i < n2; ++i) {
            const cv::KeyPoint& k = F2.mvKeysUn[i];
            kp2[i].x = k.pt.x; kp2[i].y = k.pt.y; kp2[i].size = k.size; kp2[i].angle = k.angle; kp2[i].response = k.response; kp2[i].octave = k.octave;
        }
        std::vector<float> x(n1), y(n1), r(n1);
        std::vector<int32_t> lo(n1, 0), hi(n1, 0);   // GetFeaturesInArea(..., minLevel 0, maxLevel 0)
        for (size_t i = 0; i < n1; ++i) {
            x[i] = centres[i].x; y[i] = centres[i].y;
            r[i] = F1.mvKeysUn[i].octave <= 0 ? (float)windowSize : 0.f;   // radius 0: an empty window, like the skipped keypoints (431-432)
        }
        const float bounds[4] = {(float)FrameT::mnMinX, (float)FrameT::mnMinY, (float)FrameT::mnMaxX, (float)FrameT::mnMaxY};
        const std::vector<uint8_t> A = pack(F1.mDescriptors), B = pack(F2.mDescriptors);
        int cap = (int)std::max<size_t>(4096, 32 * n1);
        for (int attempt = 0; attempt < 2; ++attempt) {
            cands.resize((size_t)cap); dist.resize((size_t)cap);
            int32_t total = 0;
            const int rc = orbm_window_lists(device_, kp2.data(), (int)n2, bounds, B.data(), A.data(), (int)n1, x.data(), y.data(), r.data(), lo.data(),
                                             hi.data(), offsets.data(), cands.data(), dist.data(), cap, &total);
            if (rc == ORB_OK) { cands.resize((size_t)total); dist.resize((size_t)total); return true; }
            if (rc != ORB_ECAPACITY) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error());
            cap = total;   // the lists need this much room: once more
        }
        throw std::runtime_error("orb_b200: window lists did not fit their own size");
    }
    template <class FrameT>
    typename std::enable_if<!b200_detail::has_image_bounds<FrameT>::value, bool>::type
    window_lists_on_device(FrameT&, FrameT&, const std::vector<cv::Point2f>&, int, std::vector<int32_t>&, std::vector<int32_t>&, std::vector<int16_t>&) {
        return false;
    }
    static std::vector<uint8_t> pack(const cv::Mat& D) {  // rows contiguous, 32 bytes each
        std::vector<uint8_t> v((size_t)D.rows * 32);
        for (int i = 0; i < D.rows; ++i) std::memcpy(&v[(size_t)i * 32], D.ptr(i), 32);
        return v;
    }
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {  // 1603-1644
        int max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < L; i++) {
            const int s = (int)histo[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
    }

    float mfNNratio;
    bool mbCheckOrientation;
    int device_;
};

}  // namespace ORB_SLAM2

#endif
