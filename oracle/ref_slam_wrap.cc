// TEST INFRASTRUCTURE ONLY (oracle). C entry points around the reference's OWN matcher and map data model:
// /root/reference/src/{ORBmatcher,Frame,KeyFrame,MapPoint,Map,KeyFrameDatabase,ORBextractor}.cc and the vendored DBoW2 are
// compiled UNMODIFIED where they lie (oracle/build_ref.sh) against the OpenCV stand-in of oracle/shim; this file only
// builds the reference's objects from plain arrays, calls the reference's methods and copies their results out.
// It pins oracle/oracle_lib.py's restatements (tests/test_oracle_vs_reference_matcher.py) and generates tests/golden.
//
// Everything below works on a "world": one Map, one KeyFrameDatabase, one vocabulary, and index-addressed Frames,
// KeyFrames and MapPoints. System* is only ever a map key in the reference (MapPoint::mbTrackInView[pSystem] ...), so
// small integers stand in for it. slam_preamble.h (force-included) turns `private` into `public` for the reference
// headers so that members can be filled from arrays; no reference line is changed.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <new>
#include <set>
#include <vector>

#include "Frame.h"
#include "KeyFrame.h"
#include "KeyFrameDatabase.h"
#include "Map.h"
#include "MapPoint.h"
#include "ORBmatcher.h"
#include "ORBextractor.h"

using namespace ORB_SLAM2;

namespace {

struct World {
    ORBVocabulary voc;
    bool has_voc = false;
    Map* map = nullptr;
    KeyFrameDatabase* db = nullptr;
    std::vector<Frame*> frames;
    std::vector<KeyFrame*> kfs;
    std::vector<MapPoint*> mps;
    std::map<MapPoint*, int> mp_id;
    std::map<KeyFrame*, int> kf_id;
    std::vector<ORBextractor*> extractors;
    ~World() {
        for (auto f : frames) delete f;
        for (auto k : kfs) delete k;
        for (auto m : mps) delete m;
        for (auto e : extractors) delete e;
        delete db; delete map;
    }
    int id_of(MapPoint* p) const { if (!p) return -1; auto it = mp_id.find(p); return it == mp_id.end() ? -2 : it->second; }
    std::vector<MapPoint*> list(const int* ids, int n) const {
        std::vector<MapPoint*> v(n);
        for (int i = 0; i < n; ++i) v[i] = ids[i] < 0 ? nullptr : mps[ids[i]];
        return v;
    }
};
inline System* sys_token(int s) { return reinterpret_cast<System*>((uintptr_t)(s + 1) * 64); }

cv::Mat mat_from(const float* p, int r, int c) {
    cv::Mat m(r, c, CV_32F);
    for (int y = 0; y < r; ++y) for (int x = 0; x < c; ++x) m.at<float>(y, x) = p[y * c + x];
    return m;
}
cv::Mat K_from(const float* k4) {
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = k4[0]; K.at<float>(1, 1) = k4[1]; K.at<float>(0, 2) = k4[2]; K.at<float>(1, 2) = k4[3];
    return K;
}
cv::Mat dist_from(const float* d, int n) {
    cv::Mat D(n < 4 ? 4 : n, 1, CV_32F);
    for (int i = 0; i < D.rows; ++i) D.at<float>(i) = i < n ? d[i] : 0.f;
    return D;
}
void fill_scale_info(Frame& F, ORBextractor* ex) {   // as every Frame ctor does, src/Frame.cc:68-75
    F.mnScaleLevels = ex->GetLevels();
    F.mfScaleFactor = ex->GetScaleFactor();
    F.mfLogScaleFactor = std::log(F.mfScaleFactor);   // float overload, as `log` under `using namespace std` there
    F.mvScaleFactors = ex->GetScaleFactors();
    F.mvInvScaleFactors = ex->GetInverseScaleFactors();
    F.mvLevelSigma2 = ex->GetScaleSigmaSquares();
    F.mvInvLevelSigma2 = ex->GetInverseScaleSigmaSquares();
}

}  // namespace

extern "C" {

void* rs_create(const char* voc_text) {
    World* w = new World;
    if (voc_text && voc_text[0]) w->has_voc = w->voc.loadFromTextFile(voc_text);
    w->map = new Map();
    w->db = new KeyFrameDatabase(w->voc);
    return w;
}
void rs_destroy(void* h) { delete (World*)h; }
int rs_has_vocabulary(void* h) { return ((World*)h)->has_voc; }

// ---- frames ---------------------------------------------------------------------------------------------------
// A Frame filled from arrays in the order of the monocular constructor (src/Frame.cc:174-228), with the extraction
// replaced by the given keypoints (x, y, size, angle, response, octave) and descriptors; UndistortKeyPoints,
// ComputeImageBounds and AssignFeaturesToGrid are the reference's own. uright / depth may be null (monocular: -1).
int rs_frame_arrays(void* h, int sys, int n, const float* kps6, const uint8_t* desc, const float* uright, const float* depth,
                    const float* K4, const float* dist, int ndist, float bf, float th_depth, int width, int height,
                    float scale_factor, int nlevels, const float* Tcw) {
    World* w = (World*)h;
    ORBextractor* ex = new ORBextractor(1000, scale_factor, nlevels, 20, 7);
    w->extractors.push_back(ex);
    Frame* F = new Frame();
    F->mpSystem = sys_token(sys);
    F->mpORBvocabulary = &w->voc;
    F->mpORBextractorLeft = ex; F->mpORBextractorRight = nullptr;
    F->mTimeStamp = (double)w->frames.size();
    F->mK = K_from(K4); F->mDistCoef = dist_from(dist, ndist);
    F->mbf = bf; F->mThDepth = th_depth;
    F->mpReferenceKF = nullptr;
    F->mnId = Frame::nNextId++;
    fill_scale_info(*F, ex);
    F->mvKeys.resize(n);
    for (int i = 0; i < n; ++i) {
        const float* k = kps6 + 6 * i;
        F->mvKeys[i] = cv::KeyPoint(k[0], k[1], k[2], k[3], k[4], (int)k[5]);
    }
    F->mDescriptors = cv::Mat(n, 32, CV_8U);
    if (n) std::memcpy(F->mDescriptors.data, desc, (size_t)n * 32);
    F->N = n;
    F->UndistortKeyPoints();
    F->mvuRight.assign(n, -1.f); F->mvDepth.assign(n, -1.f);
    if (uright) F->mvuRight.assign(uright, uright + n);
    if (depth) F->mvDepth.assign(depth, depth + n);
    F->mvpMapPoints.assign(n, (MapPoint*)nullptr);
    F->mvbOutlier.assign(n, false);
    cv::Mat dummy(height, width, CV_8U);
    F->ComputeImageBounds(dummy);
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    Frame::fx = K4[0]; Frame::fy = K4[1]; Frame::cx = K4[2]; Frame::cy = K4[3];
    Frame::invfx = 1.0f / Frame::fx; Frame::invfy = 1.0f / Frame::fy;
    Frame::mbInitialComputations = false;
    F->mb = F->mbf / Frame::fx;
    F->AssignFeaturesToGrid();
    if (Tcw) F->SetPose(mat_from(Tcw, 4, 4));
    w->frames.push_back(F);
    return (int)w->frames.size() - 1;
}

// The reference's own constructors on images. kind 0: monocular (src/Frame.cc:174), 1: stereo (61; imgR), 2: RGB-D (119;
// depth = float image). The extractor(s) are the reference's ORBextractor.
int rs_frame_images(void* h, int sys, int kind, const uint8_t* img, const uint8_t* imgR, const float* depth, int width, int height,
                    const float* K4, const float* dist, int ndist, float bf, float th_depth, int nfeatures, float scale_factor,
                    int nlevels, int ini_th, int min_th, float mb_before) {
    World* w = (World*)h;
    ORBextractor* exL = new ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
    ORBextractor* exR = kind == 1 ? new ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th) : nullptr;
    w->extractors.push_back(exL);
    if (exR) w->extractors.push_back(exR);
    cv::Mat K = K_from(K4), D = dist_from(dist, ndist);
    cv::Mat im(height, width, CV_8U, (void*)img, (size_t)width);
    Frame::mbInitialComputations = true;
    Frame* F;
    const double ts = (double)w->frames.size();
    if (kind == 0) F = new Frame(sys_token(sys), im, ts, exL, &w->voc, K, D, bf, th_depth);
    else if (kind == 1) {
        cv::Mat imr(height, width, CV_8U, (void*)imgR, (size_t)width);
        // The stereo constructor runs ComputeStereoMatches (which reads mb: minZ = mb, maxD = mbf/minZ, src/Frame.cc:489-491)
        // BEFORE it assigns mb = mbf/fx (src/Frame.cc:114), and no initialiser touches mb: the reference matches with whatever
        // the object's memory held - in the running system the previous frame's baseline. The object is therefore built in
        // memory that already holds `mb_before`, which makes that quirk an explicit input.
        void* mem = ::operator new(sizeof(Frame));
        std::memset(mem, 0, sizeof(Frame));
        *reinterpret_cast<float*>(static_cast<char*>(mem) + offsetof(Frame, mb)) = mb_before;
        F = new (mem) Frame(sys_token(sys), im, imr, ts, exL, exR, &w->voc, K, D, bf, th_depth);
    } else {
        cv::Mat dm(height, width, CV_32F, (void*)depth, (size_t)width * 4);
        F = new Frame(sys_token(sys), im, dm, ts, exL, &w->voc, K, D, bf, th_depth);
    }
    w->frames.push_back(F);
    return (int)w->frames.size() - 1;
}
int rs_frame_n(void* h, int f) { return ((World*)h)->frames[f]->N; }
void rs_frame_get(void* h, int f, float* kps6, float* un2, uint8_t* desc, float* uright, float* depth) {
    Frame& F = *((World*)h)->frames[f];
    for (int i = 0; i < F.N; ++i) {
        const cv::KeyPoint& k = F.mvKeys[i];
        const float rec[6] = {k.pt.x, k.pt.y, k.size, k.angle, k.response, (float)k.octave};
        if (kps6) std::memcpy(kps6 + 6 * i, rec, sizeof(rec));
        if (un2) { un2[2 * i] = F.mvKeysUn[i].pt.x; un2[2 * i + 1] = F.mvKeysUn[i].pt.y; }
        if (desc) std::memcpy(desc + 32 * i, F.mDescriptors.ptr(i), 32);
        if (uright) uright[i] = F.mvuRight[i];
        if (depth) depth[i] = F.mvDepth[i];
    }
}
void rs_frame_bounds(void* h, int f, float* out6) {
    (void)h; (void)f;
    out6[0] = Frame::mnMinX; out6[1] = Frame::mnMaxX; out6[2] = Frame::mnMinY; out6[3] = Frame::mnMaxY;
    out6[4] = Frame::mfGridElementWidthInv; out6[5] = Frame::mfGridElementHeightInv;
}
void rs_frame_set_pose(void* h, int f, const float* Tcw) { ((World*)h)->frames[f]->SetPose(mat_from(Tcw, 4, 4)); }
void rs_frame_set_mappoints(void* h, int f, const int* mp) {
    World* w = (World*)h; Frame& F = *w->frames[f];
    for (int i = 0; i < F.N; ++i) F.mvpMapPoints[i] = mp[i] < 0 ? nullptr : w->mps[mp[i]];
}
void rs_frame_set_outliers(void* h, int f, const uint8_t* flags) {
    Frame& F = *((World*)h)->frames[f];
    for (int i = 0; i < F.N; ++i) F.mvbOutlier[i] = flags[i] != 0;
}
void rs_frame_get_mappoints(void* h, int f, int* mp) {
    World* w = (World*)h; Frame& F = *w->frames[f];
    for (int i = 0; i < F.N; ++i) mp[i] = w->id_of(F.mvpMapPoints[i]);
}
// the 64 x 48 grid as built by Frame::AssignFeaturesToGrid: counts[64*48] (ix-major) and the items in cell order
int rs_frame_grid(void* h, int f, int* counts, int* items) {
    Frame& F = *((World*)h)->frames[f];
    int n = 0;
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) {
            counts[i * FRAME_GRID_ROWS + j] = (int)F.mGrid[i][j].size();
            for (size_t k : F.mGrid[i][j]) items[n++] = (int)k;
        }
    return n;
}
int rs_frame_features_in_area(void* h, int f, float x, float y, float r, int min_level, int max_level, int* out, int cap) {
    std::vector<size_t> v = ((World*)h)->frames[f]->GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}
int rs_kf_features_in_area(void* h, int k, float x, float y, float r, int* out, int cap) {
    std::vector<size_t> v = ((World*)h)->kfs[k]->GetFeaturesInArea(x, y, r);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}

// raw object pointers, for tests/cpp/real_types_test.cc (the product's facade templates instantiated on these types)
void* rs_frame_ptr(void* h, int f) { return ((World*)h)->frames[f]; }
void* rs_kf_ptr(void* h, int k) { return ((World*)h)->kfs[k]; }
void* rs_mp_ptr(void* h, int m) { return ((World*)h)->mps[m]; }
int rs_mp_id(void* h, void* p) { return ((World*)h)->id_of((MapPoint*)p); }

// ---- bag of words ----------------------------------------------------------------------------------------------
static int dump_bow(const DBoW2::BowVector& bv, const DBoW2::FeatureVector& fv, int* word_ids, double* word_w, int* node_ids,
                    int* node_off, int* feat, int* n_nodes) {
    int i = 0;
    for (auto& kv : bv) { if (word_ids) { word_ids[i] = (int)kv.first; word_w[i] = kv.second; } ++i; }
    int nn = 0, nf = 0;
    for (auto& kv : fv) {
        if (node_ids) { node_ids[nn] = (int)kv.first; node_off[nn] = nf; for (unsigned v : kv.second) feat[nf++] = (int)v; }
        else nf += (int)kv.second.size();
        ++nn;
    }
    if (node_off) node_off[nn] = nf;
    *n_nodes = nn;
    return i;
}
int rs_frame_compute_bow(void* h, int f, int* word_ids, double* word_w, int* node_ids, int* node_off, int* feat, int* n_nodes) {
    Frame& F = *((World*)h)->frames[f];
    F.ComputeBoW();
    return dump_bow(F.mBowVec, F.mFeatVec, word_ids, word_w, node_ids, node_off, feat, n_nodes);
}
int rs_kf_compute_bow(void* h, int k, int* word_ids, double* word_w, int* node_ids, int* node_off, int* feat, int* n_nodes) {
    KeyFrame& K = *((World*)h)->kfs[k];
    K.ComputeBoW();
    return dump_bow(K.mBowVec, K.mFeatVec, word_ids, word_w, node_ids, node_off, feat, n_nodes);
}
static void fill_featvec(DBoW2::FeatureVector& fv, int n_nodes, const int* node_ids, const int* node_off, const int* feat) {
    fv.clear();
    for (int i = 0; i < n_nodes; ++i)
        for (int j = node_off[i]; j < node_off[i + 1]; ++j) fv.addFeature((DBoW2::NodeId)node_ids[i], (unsigned)feat[j]);
}
void rs_frame_set_featvec(void* h, int f, int n_nodes, const int* node_ids, const int* node_off, const int* feat) {
    fill_featvec(((World*)h)->frames[f]->mFeatVec, n_nodes, node_ids, node_off, feat);
}
void rs_kf_set_featvec(void* h, int k, int n_nodes, const int* node_ids, const int* node_off, const int* feat) {
    fill_featvec(((World*)h)->kfs[k]->mFeatVec, n_nodes, node_ids, node_off, feat);
}
void rs_frame_set_bowvec(void* h, int f, int n, const int* ids, const double* wt) {
    DBoW2::BowVector& bv = ((World*)h)->frames[f]->mBowVec;
    bv.clear();
    for (int i = 0; i < n; ++i) bv.addWeight((DBoW2::WordId)ids[i], wt[i]);
}
// mCovisScore is read but never written by DetectCovisibilityCandidates (src/KeyFrameDatabase.cc:270-276) and mRelocScore of a
// keyframe below the common-word floor is read stale (378-387); neither is initialised by the KeyFrame constructor. The
// tests make both explicit inputs / outputs.
void rs_kf_set_scores(void* h, int k, float covis, float reloc) { KeyFrame& K = *((World*)h)->kfs[k]; K.mCovisScore = covis; K.mRelocScore = reloc; }
float rs_kf_reloc_score(void* h, int k) { return ((World*)h)->kfs[k]->mRelocScore; }
void rs_kf_set_bowvec(void* h, int k, int n, const int* ids, const double* wt) {
    DBoW2::BowVector& bv = ((World*)h)->kfs[k]->mBowVec;
    bv.clear();
    for (int i = 0; i < n; ++i) bv.addWeight((DBoW2::WordId)ids[i], wt[i]);
}

// ---- keyframes and map points --------------------------------------------------------------------------------------
int rs_keyframe(void* h, int f, int sys) {
    World* w = (World*)h;
    Frame& F = *w->frames[f];
    if (F.mTcw.empty()) F.SetPose(cv::Mat::eye(4, 4, CV_32F));
    KeyFrame* k = new KeyFrame(F, w->map, w->db, sys_token(sys));
    w->map->AddKeyFrame(k);
    w->kf_id[k] = (int)w->kfs.size();
    w->kfs.push_back(k);
    return (int)w->kfs.size() - 1;
}
void rs_kf_set_pose(void* h, int k, const float* Tcw) { ((World*)h)->kfs[k]->SetPose(mat_from(Tcw, 4, 4)); }
void rs_kf_get_mappoints(void* h, int k, int* mp) {
    World* w = (World*)h;
    std::vector<MapPoint*> v = w->kfs[k]->GetMapPointMatches();
    for (size_t i = 0; i < v.size(); ++i) mp[i] = w->id_of(v[i]);
}
void rs_kf_set_mappoints(void* h, int k, const int* mp) {   // slots only, no observation bookkeeping
    World* w = (World*)h; KeyFrame& K = *w->kfs[k];
    for (int i = 0; i < K.N; ++i) K.mvpMapPoints[i] = mp[i] < 0 ? nullptr : w->mps[mp[i]];
}
void rs_kf_add_connection(void* h, int a, int b, int weight) { ((World*)h)->kfs[a]->AddConnection(((World*)h)->kfs[b], weight); }

int rs_mappoint(void* h, const float* pos3, int ref_kf) {
    World* w = (World*)h;
    MapPoint* p = new MapPoint(mat_from(pos3, 3, 1), w->kfs[ref_kf], w->map);
    w->map->AddMapPoint(p);
    w->mp_id[p] = (int)w->mps.size();
    w->mps.push_back(p);
    return (int)w->mps.size() - 1;
}
// MapPoint(Pos, pMap, pFrame, idxF), src/MapPoint.cc:48-72 (normal, distance range and descriptor from the frame)
int rs_mappoint_from_frame(void* h, const float* pos3, int f, int idx) {
    World* w = (World*)h;
    MapPoint* p = new MapPoint(mat_from(pos3, 3, 1), w->map, w->frames[f], idx);
    w->mp_id[p] = (int)w->mps.size();
    w->mps.push_back(p);
    return (int)w->mps.size() - 1;
}
// direct state: any pointer may be null (left as is)
void rs_mp_set(void* h, int m, const uint8_t* desc32, const float* normal3, const float* min_max_dist, const int* n_obs, const int* bad) {
    MapPoint& p = *((World*)h)->mps[m];
    if (desc32) { p.mDescriptor = cv::Mat(1, 32, CV_8U); std::memcpy(p.mDescriptor.data, desc32, 32); }
    if (normal3) p.mNormalVector = mat_from(normal3, 3, 1);
    if (min_max_dist) { p.mfMinDistance = min_max_dist[0]; p.mfMaxDistance = min_max_dist[1]; }
    if (n_obs) p.nObs = *n_obs;
    if (bad) p.mbBad = *bad != 0;
}
void rs_mp_get(void* h, int m, float* pos3, uint8_t* desc32, float* normal3, float* min_max_dist, int* n_obs, int* bad) {
    MapPoint& p = *((World*)h)->mps[m];
    for (int i = 0; i < 3; ++i) { if (pos3) pos3[i] = p.mWorldPos.at<float>(i); if (normal3) normal3[i] = p.mNormalVector.at<float>(i); }
    if (desc32 && !p.mDescriptor.empty()) std::memcpy(desc32, p.mDescriptor.data, 32);
    if (min_max_dist) { min_max_dist[0] = p.mfMinDistance; min_max_dist[1] = p.mfMaxDistance; }
    if (n_obs) *n_obs = p.nObs;
    if (bad) *bad = p.mbBad;
}
// observation both ways, as the reference's callers do (e.g. src/LocalMapping.cc:432-436)
void rs_observe(void* h, int m, int k, int idx) {
    World* w = (World*)h;
    w->mps[m]->AddObservation(w->kfs[k], idx);
    w->kfs[k]->AddMapPoint(w->mps[m], idx);
}
// mObservations[kf] = idx without touching nObs (so that a test can choose the observation counts freely)
void rs_mp_set_observation(void* h, int m, int k, int idx) { World* w = (World*)h; w->mps[m]->mObservations[w->kfs[k]] = (size_t)idx; }
void rs_mp_compute_distinctive(void* h, int m) { ((World*)h)->mps[m]->ComputeDistinctiveDescriptors(); }
void rs_mp_update_normal_and_depth(void* h, int m) { ((World*)h)->mps[m]->UpdateNormalAndDepth(); }
void rs_mp_set_track(void* h, int m, int sys, int in_view, float px, float py, float pxr, int level, float view_cos) {
    MapPoint& p = *((World*)h)->mps[m];
    System* s = sys_token(sys);
    p.mbTrackInView[s] = in_view != 0; p.mTrackProjX[s] = px; p.mTrackProjY[s] = py; p.mTrackProjXR[s] = pxr;
    p.mnTrackScaleLevel[s] = level; p.mTrackViewCos[s] = view_cos;
}
// Frame::isInFrustum, src/Frame.cc:269-325; out: in_view, u, v, ur, level, view_cos
int rs_is_in_frustum(void* h, int f, int m, float cos_limit, float* out6) {
    World* w = (World*)h;
    Frame& F = *w->frames[f]; MapPoint* p = w->mps[m];
    const bool r = F.isInFrustum(p, cos_limit, F.mpSystem);
    System* s = F.mpSystem;
    out6[0] = r; out6[1] = p->mTrackProjX[s]; out6[2] = p->mTrackProjY[s]; out6[3] = p->mTrackProjXR[s];
    out6[4] = (float)p->mnTrackScaleLevel[s]; out6[5] = p->mTrackViewCos[s];
    return r;
}
int rs_predict_scale(void* h, int m, float dist, int f) { World* w = (World*)h; return w->mps[m]->PredictScale(dist, w->frames[f]); }

// ---- the eleven searches of include/ORBmatcher.h:37-102 ------------------------------------------------------------
int rs_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma(1, 32, CV_8U, (void*)a, 32), mb(1, 32, CV_8U, (void*)b, 32);
    return ORBmatcher::DescriptorDistance(ma, mb);
}
// SearchByProjection(Frame&, vector<MapPoint*>&, th), src/ORBmatcher.cc:45-131. Result: rs_frame_get_mappoints.
int rs_search_by_projection_local(void* h, float nnratio, int f, const int* mp_ids, int n, float th) {
    World* w = (World*)h;
    std::vector<MapPoint*> v = w->list(mp_ids, n);
    return ORBmatcher(nnratio, true).SearchByProjection(*w->frames[f], v, th);
}
// SearchByProjection(Cur, Last, th, mono), 1330-1472
int rs_search_by_projection_last(void* h, float nnratio, int check_ori, int cur, int last, float th, int mono) {
    World* w = (World*)h;
    return ORBmatcher(nnratio, check_ori != 0).SearchByProjection(*w->frames[cur], *w->frames[last], th, mono != 0);
}
// SearchByProjection(Cur, KF, sAlreadyFound, th, ORBdist), 1474-1601
int rs_search_by_projection_kf(void* h, float nnratio, int check_ori, int cur, int kf, const int* found, int n_found, float th, int orb_dist) {
    World* w = (World*)h;
    std::set<MapPoint*> s;
    for (int i = 0; i < n_found; ++i) s.insert(w->mps[found[i]]);
    return ORBmatcher(nnratio, check_ori != 0).SearchByProjection(*w->frames[cur], w->kfs[kf], s, th, orb_dist);
}
// SearchByProjection(KF, Scw, vpPoints, vpMatched, th), 292-405; matched[N] in/out as map point ids
int rs_search_by_projection_sim3(void* h, float nnratio, int kf, const float* Scw, const int* pts, int n, int* matched, int th) {
    World* w = (World*)h;
    KeyFrame* k = w->kfs[kf];
    std::vector<MapPoint*> v = w->list(pts, n), m = w->list(matched, k->N);
    const int r = ORBmatcher(nnratio, true).SearchByProjection(k, mat_from(Scw, 4, 4), v, m, th);
    for (int i = 0; i < k->N; ++i) matched[i] = w->id_of(m[i]);
    return r;
}
// SearchByBoW(KF, F, vpMapPointMatches), 161-290; out[F.N]
int rs_search_by_bow_kf_f(void* h, float nnratio, int check_ori, int kf, int f, int* out) {
    World* w = (World*)h;
    std::vector<MapPoint*> m;
    const int r = ORBmatcher(nnratio, check_ori != 0).SearchByBoW(w->kfs[kf], *w->frames[f], m);
    for (size_t i = 0; i < m.size(); ++i) out[i] = w->id_of(m[i]);
    return r;
}
// SearchByBoW(KF1, KF2, vpMatches12), 524-657; out[KF1.N]
int rs_search_by_bow_kf_kf(void* h, float nnratio, int check_ori, int k1, int k2, int* out) {
    World* w = (World*)h;
    std::vector<MapPoint*> m;
    const int r = ORBmatcher(nnratio, check_ori != 0).SearchByBoW(w->kfs[k1], w->kfs[k2], m);
    for (size_t i = 0; i < m.size(); ++i) out[i] = w->id_of(m[i]);
    return r;
}
// SearchForInitialization, 407-522; prev_matched[2*N1] in/out, matches12[N1] out
int rs_search_for_initialization(void* h, float nnratio, int check_ori, int f1, int f2, float* prev_matched, int* matches12, int window) {
    World* w = (World*)h;
    Frame& F1 = *w->frames[f1];
    std::vector<cv::Point2f> pm(F1.N);
    for (int i = 0; i < F1.N; ++i) pm[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> m12;
    const int r = ORBmatcher(nnratio, check_ori != 0).SearchForInitialization(F1, *w->frames[f2], pm, m12, window);
    for (int i = 0; i < F1.N; ++i) { prev_matched[2 * i] = pm[i].x; prev_matched[2 * i + 1] = pm[i].y; matches12[i] = m12[i]; }
    return r;
}
// SearchForTriangulation, 659-825; pairs out (2 ints each), returns the number of pairs
int rs_search_for_triangulation(void* h, float nnratio, int check_ori, int k1, int k2, const float* F12, int only_stereo, int* pairs, int cap) {
    World* w = (World*)h;
    std::vector<std::pair<size_t, size_t> > v;
    ORBmatcher(nnratio, check_ori != 0).SearchForTriangulation(w->kfs[k1], w->kfs[k2], mat_from(F12, 3, 3), v, only_stereo != 0);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) { pairs[2 * i] = (int)v[i].first; pairs[2 * i + 1] = (int)v[i].second; }
    return (int)v.size();
}
// SearchBySim3, 1104-1328; matches12[KF1.N] in/out as map point ids of KF2
int rs_search_by_sim3(void* h, float nnratio, int k1, int k2, int* matches12, float s12, const float* R12, const float* t12, float th) {
    World* w = (World*)h;
    KeyFrame* a = w->kfs[k1];
    std::vector<MapPoint*> m = w->list(matches12, a->N);
    const int r = ORBmatcher(nnratio, true).SearchBySim3(a, w->kfs[k2], m, s12, mat_from(R12, 3, 3), mat_from(t12, 3, 1), th);
    for (int i = 0; i < a->N; ++i) matches12[i] = w->id_of(m[i]);
    return r;
}
// Fuse(KF, vpMapPoints, th), 827-977. The graph edits it performs (AddObservation / Replace) are the reference's; the
// resulting keyframe slots are read back with rs_kf_get_mappoints and per-point state with rs_mp_get / rs_mp_replaced.
int rs_fuse(void* h, float nnratio, int kf, const int* pts, int n, float th) {
    World* w = (World*)h;
    std::vector<MapPoint*> v = w->list(pts, n);
    return ORBmatcher(nnratio, true).Fuse(w->kfs[kf], v, th);
}
// Fuse(KF, Scw, vpPoints, th, vpReplacePoint), 979-1102; replace[n] in/out as map point ids
int rs_fuse_sim3(void* h, float nnratio, int kf, const float* Scw, const int* pts, int n, float th, int* replace) {
    World* w = (World*)h;
    std::vector<MapPoint*> v = w->list(pts, n), rp = w->list(replace, n);
    const int r = ORBmatcher(nnratio, true).Fuse(w->kfs[kf], mat_from(Scw, 4, 4), v, th, rp);
    for (int i = 0; i < n; ++i) replace[i] = w->id_of(rp[i]);
    return r;
}
int rs_mp_replaced(void* h, int m) { World* w = (World*)h; return w->id_of(w->mps[m]->GetReplaced()); }
int rs_mp_index_in_kf(void* h, int m, int k) { World* w = (World*)h; return w->mps[m]->GetIndexInKeyFrame(w->kfs[k]); }

// ---- keyframe database (src/KeyFrameDatabase.cc) -------------------------------------------------------------------
void rs_db_add(void* h, int k) { World* w = (World*)h; w->db->add(w->kfs[k]); }
void rs_db_erase(void* h, int k) { World* w = (World*)h; w->db->erase(w->kfs[k]); }
int rs_db_detect_loop_candidates(void* h, int k, float min_score, int* out, int cap) {
    World* w = (World*)h;
    std::vector<KeyFrame*> v = w->db->DetectLoopCandidates(w->kfs[k], min_score);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = w->kf_id[v[i]];
    return (int)v.size();
}
int rs_db_detect_relocalization_candidates(void* h, int f, int* out, int cap) {
    World* w = (World*)h;
    std::vector<KeyFrame*> v = w->db->DetectRelocalizationCandidates(w->frames[f]);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = w->kf_id[v[i]];
    return (int)v.size();
}
int rs_db_detect_covisibility_candidates(void* h, int k, float min_score, const int* ignore, int n_ignore, int* out, int cap) {
    World* w = (World*)h;
    std::vector<KeyFrame*> ig;
    for (int i = 0; i < n_ignore; ++i) ig.push_back(w->kfs[ignore[i]]);
    std::vector<KeyFrame*> v = w->db->DetectCovisibilityCandidates(w->kfs[k], min_score, ig);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = w->kf_id[v[i]];
    return (int)v.size();
}

// ---- the float cv::Mat arithmetic of the stand-in, exported so that it can be pinned to cv2 (tests/test_oracle_vs_cv2.py)
void rs_gemm32f(const float* A, int ar, int ac, const float* B, int br, int bc, double alpha, const float* C, double beta, float* D, int flags) {
    const int n = (flags & cvprim::GEMM_B_T) ? br : bc;
    cvprim::gemm32f(A, ar, ac, ac, B, br, bc, bc, alpha, C, n, beta, D, n, flags);
}
double rs_norm_l2(const float* p, int n) { return cvprim::norm_l2_32f(p, n, 1, 1); }
double rs_dot(const float* a, const float* b, int n) { return cvprim::dot_32f(a, 1, b, 1, n, 1); }
void rs_undistort_points(const float* src, float* dst, int n, const float* K9, const float* dist, int nd) {
    cvprim::undistort_points_32f(src, dst, n, K9, 3, dist, nd);
}

}  // extern "C"
