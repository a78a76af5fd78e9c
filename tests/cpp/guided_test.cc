// Test driver for the guided searches of the C++ facade (include/orbslam2_b200/ORBmatcher.h): reads a scenario written by
// tests/test_gpu_cpp_guided.py (two views with keypoints / descriptors / poses / feature vectors, a set of map points),
// builds stand-ins with the members the reference's Frame / KeyFrame / MapPoint expose to ORBmatcher, runs the nine
// searches and dumps their results for the Python side to compare with the oracle.
//   guided_test <scenario.bin> <out.bin>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#include "orbslam2_b200/FrameGrid.h"
#include "orbslam2_b200/ORBmatcher.h"

struct SystemLite {};
struct KeyFrameLite;

static std::vector<int> g_log;  // graph operations performed by Fuse, for the comparison

struct MapPointLite {
    int id = 0;
    bool bad = false;
    cv::Mat pos, normal, desc;
    float maxD = 0, minD = 0;
    int nObs = 0;
    std::map<SystemLite*, bool> mbTrackInView;
    std::map<SystemLite*, int> mnTrackScaleLevel;
    std::map<SystemLite*, float> mTrackViewCos, mTrackProjX, mTrackProjY, mTrackProjXR;
    std::map<KeyFrameLite*, size_t> obs;
    bool isBad() const { return bad; }
    cv::Mat GetWorldPos() const { return pos; }
    cv::Mat GetNormal() const { return normal; }
    cv::Mat GetDescriptor() const { return desc; }
    float GetMinDistanceInvariance() const { return 0.8f * minD; }
    float GetMaxDistanceInvariance() const { return 1.2f * maxD; }
    int Observations() const { return nObs; }
    template <class T> int PredictScale(const float& currentDist, T* pF) const {  // src/MapPoint.cc:389-421
        const float ratio = maxD / currentDist;
        int nScale = (int)std::ceil(std::log(ratio) / pF->mfLogScaleFactor);
        if (nScale < 0) nScale = 0;
        else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
        return nScale;
    }
    bool IsInKeyFrame(KeyFrameLite* kf) const { return obs.count(kf) != 0; }
    int GetIndexInKeyFrame(KeyFrameLite* kf) const { auto it = obs.find(kf); return it == obs.end() ? -1 : (int)it->second; }
    void AddObservation(KeyFrameLite* kf, size_t idx) { g_log.push_back(1); g_log.push_back(id); g_log.push_back((int)idx); obs[kf] = idx; nObs++; }
    void Replace(MapPointLite* other) { g_log.push_back(2); g_log.push_back(id); g_log.push_back(other->id); bad = true; }
};

struct ViewLite {  // members shared by the Frame and KeyFrame stand-ins
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    cv::Mat mDescriptors, mTcw;
    std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<MapPointLite*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::map<unsigned, std::vector<unsigned> > mFeatVec;
    float fx, fy, cx, cy, mbf, mb, mnMinX, mnMaxX, mnMinY, mnMaxY, mfLogScaleFactor;
    int mnScaleLevels;
    SystemLite* mpSystem = nullptr;
    ORB_SLAM2::FrameGrid<cv::KeyPoint> grid;
};
struct FrameLite : ViewLite {
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
        return grid.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    }
};
struct KeyFrameLite : ViewLite {
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const { return grid.GetFeaturesInArea(x, y, r); }
    std::vector<MapPointLite*> GetMapPointMatches() { return mvpMapPoints; }
    MapPointLite* GetMapPoint(size_t idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPointLite* p, size_t idx) { mvpMapPoints[idx] = p; }
    std::set<MapPointLite*> GetMapPoints() {
        std::set<MapPointLite*> s;
        for (size_t i = 0; i < mvpMapPoints.size(); ++i)
            if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
        return s;
    }
    bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }
    cv::Mat GetRotation() const { cv::Mat R(3, 3, CV_32F); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.at<float>(i, j) = mTcw.at<float>(i, j); return R; }
    cv::Mat GetTranslation() const { cv::Mat t(3, 1, CV_32F); for (int i = 0; i < 3; ++i) t.at<float>(i, 0) = mTcw.at<float>(i, 3); return t; }
    cv::Mat Ow;
    cv::Mat GetCameraCenter() const { return Ow; }
};

static FILE* g_in;
template <class T> static std::vector<T> rd(size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, g_in) != n) { fprintf(stderr, "scenario truncated\n"); exit(2); }
    return v;
}
static cv::Mat matf(int r, int c, const float* p) { cv::Mat m(r, c, CV_32F); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = p[i * c + j]; return m; }

static FILE* g_out;
static void wr(const std::vector<int>& v) { const int n = (int)v.size(); fwrite(&n, 4, 1, g_out); fwrite(v.data(), 4, v.size(), g_out); }
static std::vector<int> ids(const std::vector<MapPointLite*>& v) { std::vector<int> r(v.size()); for (size_t i = 0; i < v.size(); ++i) r[i] = v[i] ? v[i]->id : -1; return r; }

template <class V> static void load_view(V& f, const std::vector<float>& cam, const std::vector<float>& scale, SystemLite* sys) {
    f.N = rd<int>(1)[0];
    const std::vector<float> k = rd<float>((size_t)f.N * 6);
    f.mvKeys.resize(f.N);
    for (int i = 0; i < f.N; ++i) {
        cv::KeyPoint& kp = f.mvKeys[i];
        kp.pt.x = k[6 * i]; kp.pt.y = k[6 * i + 1]; kp.size = k[6 * i + 2]; kp.angle = k[6 * i + 3]; kp.response = k[6 * i + 4]; kp.octave = (int)k[6 * i + 5];
    }
    f.mvKeysUn = f.mvKeys;
    const std::vector<unsigned char> d = rd<unsigned char>((size_t)f.N * 32);
    f.mDescriptors.create(f.N, 32, CV_8U);
    for (int i = 0; i < f.N; ++i) memcpy(f.mDescriptors.ptr(i), &d[(size_t)i * 32], 32);
    f.mvuRight = rd<float>(f.N);
    f.mTcw = matf(4, 4, rd<float>(16).data());
    const int nn = rd<int>(1)[0];
    for (int n = 0; n < nn; ++n) {
        const std::vector<int> h = rd<int>(2);
        const std::vector<int> idx = rd<int>(h[1]);
        f.mFeatVec[(unsigned)h[0]].assign(idx.begin(), idx.end());
    }
    f.fx = cam[0]; f.fy = cam[1]; f.cx = cam[2]; f.cy = cam[3]; f.mbf = cam[4]; f.mb = cam[5];
    f.mnMinX = cam[6]; f.mnMaxX = cam[7]; f.mnMinY = cam[8]; f.mnMaxY = cam[9]; f.mfLogScaleFactor = cam[10];
    f.mnScaleLevels = (int)scale.size();
    f.mvScaleFactors = scale;
    for (size_t l = 0; l < scale.size(); ++l) { f.mvLevelSigma2.push_back(scale[l] * scale[l]); f.mvInvLevelSigma2.push_back(1.0f / (scale[l] * scale[l])); }
    f.mpSystem = sys;
    f.grid.SetBounds(f.mnMinX, f.mnMinY, f.mnMaxX, f.mnMaxY);
    f.grid.Assign(f.mvKeysUn);
    f.mvpMapPoints.assign(f.N, (MapPointLite*)NULL);
    f.mvbOutlier.assign(f.N, false);
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    g_in = fopen(argv[1], "rb");
    g_out = fopen(argv[2], "wb");
    if (!g_in || !g_out) return 2;
    SystemLite sys;
    const int nl = rd<int>(1)[0];
    const std::vector<float> scale = rd<float>(nl), cam = rd<float>(11);
    FrameLite F[2];
    KeyFrameLite K[2];
    // the same two views as Frame and as KeyFrame stand-ins
    const long view_start = ftell(g_in);
    for (int k = 0; k < 2; ++k) load_view(F[k], cam, scale, &sys);
    fseek(g_in, view_start, SEEK_SET);
    for (int k = 0; k < 2; ++k) {
        load_view(K[k], cam, scale, &sys);
        // mOw = -mRcw.t()*mtcw as the pose update computes it: double accumulation (see b200_detail::minusRtT)
        const ORB_SLAM2::b200_detail::Vec3 ow = ORB_SLAM2::b200_detail::minusRtT(ORB_SLAM2::b200_detail::rotation3(K[k].mTcw), ORB_SLAM2::b200_detail::column3(K[k].mTcw, 3));
        K[k].Ow = matf(3, 1, ow.v);
    }
    // map points
    const int nmp = rd<int>(1)[0];
    const std::vector<float> pos = rd<float>((size_t)nmp * 3), nrm = rd<float>((size_t)nmp * 3), maxd = rd<float>(nmp), mind = rd<float>(nmp);
    const std::vector<unsigned char> mdesc = rd<unsigned char>((size_t)nmp * 32), bad = rd<unsigned char>(nmp);
    const std::vector<int> nobs = rd<int>(nmp);
    const std::vector<unsigned char> inview = rd<unsigned char>(nmp);
    const std::vector<float> px = rd<float>(nmp), py = rd<float>(nmp), pxr = rd<float>(nmp), vcos = rd<float>(nmp);
    const std::vector<int> lvl = rd<int>(nmp);
    const std::vector<int> assoc0 = rd<int>(F[0].N), assoc1 = rd<int>(F[1].N);  // map point held by each keypoint of view 0 / 1
    const std::vector<unsigned char> outlier0 = rd<unsigned char>(F[0].N), found = rd<unsigned char>(nmp);
    const std::vector<float> F12v = rd<float>(9), Scwv = rd<float>(16), sim = rd<float>(13);  // s12, R12 (9), t12 (3)
    const std::vector<float> ths = rd<float>(8);
    fclose(g_in);

    std::vector<MapPointLite> pool;
    auto reset = [&]() {
        pool.assign(nmp, MapPointLite());
        for (int i = 0; i < nmp; ++i) {
            MapPointLite& m = pool[i];
            m.id = i; m.bad = bad[i] != 0; m.pos = matf(3, 1, &pos[3 * i]); m.normal = matf(3, 1, &nrm[3 * i]);
            m.desc.create(1, 32, CV_8U); memcpy(m.desc.ptr(), &mdesc[(size_t)i * 32], 32);
            m.maxD = maxd[i]; m.minD = mind[i]; m.nObs = nobs[i];
            m.mbTrackInView[&sys] = inview[i] != 0; m.mnTrackScaleLevel[&sys] = lvl[i]; m.mTrackViewCos[&sys] = vcos[i];
            m.mTrackProjX[&sys] = px[i]; m.mTrackProjY[&sys] = py[i]; m.mTrackProjXR[&sys] = pxr[i];
        }
        for (int k = 0; k < 2; ++k) {
            const std::vector<int>& a = k ? assoc1 : assoc0;
            for (int i = 0; i < F[k].N; ++i) {
                F[k].mvpMapPoints[i] = K[k].mvpMapPoints[i] = a[i] >= 0 ? &pool[a[i]] : (MapPointLite*)NULL;
                if (a[i] >= 0) pool[a[i]].obs[&K[k]] = (size_t)i;
            }
        }
        for (int i = 0; i < F[0].N; ++i) F[0].mvbOutlier[i] = outlier0[i] != 0;
        g_log.clear();
    };
    std::vector<MapPointLite*> all;
    auto all_points = [&]() { all.resize(nmp); for (int i = 0; i < nmp; ++i) all[i] = &pool[i]; };
    ORB_SLAM2::ORBmatcher m08(0.8f, true), m09(0.9f, true), m075(0.75f, true), m06(0.6f, false);

    // 1. SearchByProjection(F, vpMapPoints, th)                                     (a-10)
    reset(); all_points();
    for (int i = 0; i < F[1].N; ++i) if (assoc1[i] < 0) F[1].mvpMapPoints[i] = NULL;
    int n = m08.SearchByProjection(F[1], all, ths[0]);
    wr(std::vector<int>(1, n)); wr(ids(F[1].mvpMapPoints));
    // 2. SearchByProjection(Cur, Last, th, bMono) with stereo (forward / backward by the poses) and mono  (a-11)
    for (int mono = 0; mono < 2; ++mono) {
        reset();
        n = m09.SearchByProjection(F[1], F[0], ths[1], mono != 0);
        wr(std::vector<int>(1, n)); wr(ids(F[1].mvpMapPoints));
    }
    // 3. SearchByProjection(Cur, pKF, sAlreadyFound, th, ORBdist)                      (a-12)
    reset();
    std::set<MapPointLite*> sFound;
    for (int i = 0; i < nmp; ++i) if (found[i]) sFound.insert(&pool[i]);
    n = m09.SearchByProjection(F[1], &K[0], sFound, ths[2], 64);
    wr(std::vector<int>(1, n)); wr(ids(F[1].mvpMapPoints));
    // 4. SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)                         (a-12)
    reset(); all_points();
    std::vector<MapPointLite*> vpMatched = K[1].mvpMapPoints;
    n = m075.SearchByProjection(&K[1], matf(4, 4, Scwv.data()), all, vpMatched, (int)ths[3]);
    wr(std::vector<int>(1, n)); wr(ids(vpMatched));
    // 5. SearchByBoW(pKF, F, vpMapPointMatches)                                        (a-13)
    reset();
    std::vector<MapPointLite*> bow;
    n = m075.SearchByBoW(&K[0], F[1], bow);
    wr(std::vector<int>(1, n)); wr(ids(bow));
    // 6. SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)            (a-14)
    for (int only = 0; only < 2; ++only) {
        reset();
        std::vector<std::pair<size_t, size_t> > pairs;
        n = m06.SearchForTriangulation(&K[0], &K[1], matf(3, 3, F12v.data()), pairs, only != 0);
        std::vector<int> flat;
        for (size_t i = 0; i < pairs.size(); ++i) { flat.push_back((int)pairs[i].first); flat.push_back((int)pairs[i].second); }
        wr(std::vector<int>(1, n)); wr(flat);
    }
    // 7. Fuse(pKF, vpMapPoints, th)                                                    (a-15)
    reset(); all_points();
    n = m08.Fuse(&K[1], all, ths[4]);
    wr(std::vector<int>(1, n)); wr(g_log); wr(ids(K[1].mvpMapPoints));
    // 8. Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)                                  (a-15)
    reset(); all_points();
    std::vector<MapPointLite*> vpReplace(nmp, (MapPointLite*)NULL);
    n = m08.Fuse(&K[1], matf(4, 4, Scwv.data()), all, ths[5], vpReplace);
    wr(std::vector<int>(1, n)); wr(ids(vpReplace)); wr(ids(K[1].mvpMapPoints));
    // 9. SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)                      (a-15)
    reset();
    std::vector<MapPointLite*> m12(K[0].N, (MapPointLite*)NULL);
    for (int i = 0; i < K[0].N; i += 17)  // some matches known on entry: they and their keypoint in KF2 are excluded
        if (K[0].mvpMapPoints[i] && K[0].mvpMapPoints[i]->IsInKeyFrame(&K[1])) m12[i] = K[0].mvpMapPoints[i];
    const float s12 = sim[0];
    n = m08.SearchBySim3(&K[0], &K[1], m12, s12, matf(3, 3, &sim[1]), matf(3, 1, &sim[10]), ths[6]);
    wr(std::vector<int>(1, n)); wr(ids(m12));
    fclose(g_out);
    printf("guided searches ok\n");
    return 0;
}
