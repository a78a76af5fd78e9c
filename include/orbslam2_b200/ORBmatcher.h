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
#include <stdexcept>
#include <string>
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
        std::vector<int> rotHist[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
        const float factor = 1.0f / HISTO_LENGTH;
        std::vector<int> vMatchedDistance(F2.mvKeysUn.size(), INT_MAX);
        std::vector<int> vnMatches21(F2.mvKeysUn.size(), -1);

        // 1. host: gate. Candidate lists in the reference's iteration order, CSR.
        const size_t n1 = F1.mvKeysUn.size();
        std::vector<int32_t> offsets(n1 + 1, 0), cands;
        for (size_t i1 = 0; i1 < n1; i1++) {
            if (F1.mvKeysUn[i1].octave <= 0) {
                std::vector<size_t> v = F2.GetFeaturesInArea(vbPrevMatched[i1].x, vbPrevMatched[i1].y, windowSize, 0, 0);
                for (size_t k = 0; k < v.size(); ++k) cands.push_back((int32_t)v[k]);
            }
            offsets[i1 + 1] = (int32_t)cands.size();
        }
        // 2. device: all distances
        std::vector<int16_t> dist(cands.size());
        if (!cands.empty()) {
            const std::vector<uint8_t> A = pack(F1.mDescriptors), B = pack(F2.mDescriptors);
            const int rc = orbm_list_distances(device_, A.data(), (int)n1, B.data(), (int)F2.mvKeysUn.size(), offsets.data(), cands.data(), dist.data());
            if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error());
        }
        // 3. host: the reference's ordered resolve (src/ORBmatcher.cc:434-488)
        for (size_t i1 = 0; i1 < n1; i1++) {
            if (offsets[i1] == offsets[i1 + 1]) continue;
            int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
            for (int k = offsets[i1]; k < offsets[i1 + 1]; ++k) {
                const int i2 = cands[k], d = dist[k];
                if (vMatchedDistance[i2] <= d) continue;
                if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx2 = i2; }
                else if (d < bestDist2) bestDist2 = d;
            }
            if (bestDist <= TH_LOW) {
                if (bestDist < (float)bestDist2 * mfNNratio) {
                    if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                    vnMatches12[i1] = bestIdx2;
                    vnMatches21[bestIdx2] = (int)i1;
                    vMatchedDistance[bestIdx2] = bestDist;
                    nmatches++;
                    if (mbCheckOrientation) {
                        float rot = F1.mvKeysUn[i1].angle - F2.mvKeysUn[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back((int)i1);
                    }
                }
            }
        }
        if (mbCheckOrientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++) {
                if (i == ind1 || i == ind2 || i == ind3) continue;
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    const int idx1 = rotHist[i][j];
                    if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
                }
            }
        }
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
        std::vector<bool> vbMatched2(vpMapPoints2.size(), false);
        std::vector<int> rotHist[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
        const float factor = 1.0f / HISTO_LENGTH;
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
            int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
            for (int k = offsets[q]; k < offsets[q + 1]; ++k) {
                const size_t idx2 = (size_t)cands[k];
                MapPointT* pMP2 = vpMapPoints2[idx2];
                if (vbMatched2[idx2] || !pMP2) continue;
                if (pMP2->isBad()) continue;
                const int d = dist[k];
                if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdx2 = (int)idx2; }
                else if (d < bestDist2) bestDist2 = d;
            }
            if (bestDist1 < TH_LOW) {
                if (static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
                    vpMatches12[idx1] = vpMapPoints2[bestIdx2];
                    vbMatched2[bestIdx2] = true;
                    if (mbCheckOrientation) {
                        float rot = vKeysUn1[idx1].angle - vKeysUn2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back((int)idx1);
                    }
                    nmatches++;
                }
            }
        }
        if (mbCheckOrientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++) {
                if (i == ind1 || i == ind2 || i == ind3) continue;
                for (size_t j = 0; j < rotHist[i].size(); j++) { vpMatches12[rotHist[i][j]] = static_cast<MapPointT*>(NULL); nmatches--; }
            }
        }
        return nmatches;
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
