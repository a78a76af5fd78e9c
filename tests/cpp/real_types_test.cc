// Test driver: the product's C++ facade (include/orbslam2_b200/ORBmatcher.h) instantiated on the REFERENCE'S OWN Frame /
// KeyFrame / MapPoint classes (/root/reference/include/{Frame,KeyFrame,MapPoint}.h + their .cc files compiled unmodified in
// oracle/_ref/libref_slam.so), next to the reference's own ORBmatcher on identical objects. Because it needs the reference
// headers it is compiled in this container by oracle/build_ref.sh into oracle/_ref/real_types_test (a prebuilt binary
// travels to the GPU box like the other oracle/_ref outputs); tests/test_gpu_real_types.py runs it on the B200.
//
//   real_types_test <scenario.bin> <out_reference.bin> <out_b200.bin>
//
// scenario.bin is the file tests/test_gpu_cpp_guided.py::write_scenario produces. Every search runs twice, each time on a
// freshly built world: once through ORB_SLAM2::ORBmatcher (the reference, src/ORBmatcher.cc) and once through the facade
// class (renamed B200ORBmatcher here so that both fit into one translation unit). The two dumps must be identical.
#include <cstdio>
#include <cstdlib>
#include <set>
#include <string>
#include <vector>

#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "ORBmatcher.h"   // the reference's

#define ORBmatcher B200ORBmatcher
#define ORB_B200_FACADE_ON_REFERENCE_TYPES 1
#include "orbslam2_b200/ORBmatcher.h"   // the product's
#undef ORBmatcher

using namespace ORB_SLAM2;

extern "C" {   // oracle/ref_slam_wrap.cc (object construction from arrays)
void* rs_create(const char*);
void rs_destroy(void*);
int rs_frame_arrays(void*, int, int, const float*, const uint8_t*, const float*, const float*, const float*, const float*, int, float, float, int, int, float, int,
                    const float*);
int rs_keyframe(void*, int, int);
int rs_mappoint(void*, const float*, int);
void rs_mp_set(void*, int, const uint8_t*, const float*, const float*, const int*, const int*);
void rs_mp_get(void*, int, float*, uint8_t*, float*, float*, int*, int*);
void rs_frame_set_featvec(void*, int, int, const int*, const int*, const int*);
void rs_kf_set_featvec(void*, int, int, const int*, const int*, const int*);
void rs_frame_set_mappoints(void*, int, const int*);
void rs_kf_set_mappoints(void*, int, const int*);
void rs_frame_set_outliers(void*, int, const uint8_t*);
void rs_mp_set_observation(void*, int, int, int);
void* rs_frame_ptr(void*, int);
void* rs_kf_ptr(void*, int);
void* rs_mp_ptr(void*, int);
int rs_mp_id(void*, void*);
}

static FILE* g_in;
template <class T> static std::vector<T> rd(size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, g_in) != n) { fprintf(stderr, "scenario truncated\n"); exit(2); }
    return v;
}
static cv::Mat matf(int r, int c, const float* p) { cv::Mat m(r, c, CV_32F); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = p[i * c + j]; return m; }

struct View {
    int n;
    std::vector<float> k, uright, T;
    std::vector<unsigned char> d;
    std::vector<int> node_ids, node_off, feat;
};
struct Scenario {
    std::vector<float> scale, cam;
    View v[2];
    int nmp;
    std::vector<float> pos, nrm, maxd, mind;
    std::vector<unsigned char> mdesc, bad, outlier0, found;
    std::vector<int> nobs, assoc[2];
    std::vector<float> F12, Scw, sim, ths;
};

static Scenario load(const char* path) {
    Scenario s;
    g_in = fopen(path, "rb");
    if (!g_in) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    const int nl = rd<int>(1)[0];
    s.scale = rd<float>(nl); s.cam = rd<float>(11);
    for (int k = 0; k < 2; ++k) {
        View& v = s.v[k];
        v.n = rd<int>(1)[0];
        v.k = rd<float>((size_t)v.n * 6); v.d = rd<unsigned char>((size_t)v.n * 32); v.uright = rd<float>(v.n); v.T = rd<float>(16);
        const int nn = rd<int>(1)[0];
        v.node_off.push_back(0);
        for (int n = 0; n < nn; ++n) {
            const std::vector<int> h = rd<int>(2), idx = rd<int>(h[1]);
            v.node_ids.push_back(h[0]);
            v.feat.insert(v.feat.end(), idx.begin(), idx.end());
            v.node_off.push_back((int)v.feat.size());
        }
    }
    s.nmp = rd<int>(1)[0];
    const int n = s.nmp;
    s.pos = rd<float>((size_t)n * 3); s.nrm = rd<float>((size_t)n * 3); s.maxd = rd<float>(n); s.mind = rd<float>(n);
    s.mdesc = rd<unsigned char>((size_t)n * 32); s.bad = rd<unsigned char>(n); s.nobs = rd<int>(n);
    rd<unsigned char>(n); rd<float>(n); rd<float>(n); rd<float>(n); rd<float>(n); rd<int>(n);   // the oracle's frustum outputs: not used, isInFrustum runs here
    s.assoc[0] = rd<int>(s.v[0].n); s.assoc[1] = rd<int>(s.v[1].n);
    s.outlier0 = rd<unsigned char>(s.v[0].n); s.found = rd<unsigned char>(n);
    s.F12 = rd<float>(9); s.Scw = rd<float>(16); s.sim = rd<float>(13); s.ths = rd<float>(8);
    fclose(g_in);
    return s;
}

// one freshly built copy of the scenario as reference objects
struct Scene {
    void* w;
    Frame* F[2];
    KeyFrame* K[2];
    std::vector<MapPoint*> mp;
    const Scenario& s;
    explicit Scene(const Scenario& sc) : s(sc) {
        w = rs_create("");
        const float dist[4] = {0, 0, 0, 0};
        int kf[2];
        for (int k = 0; k < 2; ++k) {
            const View& v = s.v[k];
            const int f = rs_frame_arrays(w, 0, v.n, v.k.data(), v.d.data(), v.uright.data(), nullptr, s.cam.data(), dist, 4, s.cam[4], 35.f, 640, 480,
                                          s.scale[1], (int)s.scale.size(), v.T.data());
            kf[k] = rs_keyframe(w, f, 0);
            rs_frame_set_featvec(w, f, (int)v.node_ids.size(), v.node_ids.data(), v.node_off.data(), v.feat.data());
            rs_kf_set_featvec(w, kf[k], (int)v.node_ids.size(), v.node_ids.data(), v.node_off.data(), v.feat.data());
            F[k] = (Frame*)rs_frame_ptr(w, f); K[k] = (KeyFrame*)rs_kf_ptr(w, kf[k]);
        }
        for (int i = 0; i < s.nmp; ++i) {
            const int m = rs_mappoint(w, &s.pos[3 * i], kf[0]);
            const float mm[2] = {s.mind[i], s.maxd[i]};
            const int no = s.nobs[i], bd = s.bad[i];
            rs_mp_set(w, m, &s.mdesc[(size_t)i * 32], &s.nrm[3 * i], mm, &no, &bd);
            mp.push_back((MapPoint*)rs_mp_ptr(w, m));
        }
    }
    ~Scene() { rs_destroy(w); }
    void hold(int k, bool observe = false) {
        rs_frame_set_mappoints(w, k, s.assoc[k].data());
        rs_kf_set_mappoints(w, k, s.assoc[k].data());
        if (observe)
            for (int i = 0; i < s.v[k].n; ++i)
                if (s.assoc[k][i] >= 0) rs_mp_set_observation(w, s.assoc[k][i], k, i);
    }
    std::vector<int> ids(const std::vector<MapPoint*>& v) { std::vector<int> r(v.size()); for (size_t i = 0; i < v.size(); ++i) r[i] = v[i] ? rs_mp_id(w, v[i]) : -1; return r; }
};

static FILE* g_out;
static void wr(const std::vector<int>& v) { const int n = (int)v.size(); fwrite(&n, 4, 1, g_out); fwrite(v.data(), 4, v.size(), g_out); }
static void wr1(int x) { wr(std::vector<int>(1, x)); }

template <class Matcher> static void run_all(const Scenario& s, const char* out_path) {
    g_out = fopen(out_path, "wb");
    if (!g_out) exit(2);
    const std::vector<float>& th = s.ths;
    const int nmp = s.nmp;
    {   // 1. a-10: Frame::isInFrustum, then SearchByProjection(F, vpMapPoints, th)
        Scene c(s); c.hold(1);
        for (int i = 0; i < nmp; ++i) c.F[1]->isInFrustum(c.mp[i], 0.5f, c.F[1]->mpSystem);
        const int n = Matcher(0.8f, true).SearchByProjection(*c.F[1], c.mp, th[0]);
        wr1(n); wr(c.ids(c.F[1]->mvpMapPoints));
    }
    for (int mono = 0; mono < 2; ++mono) {   // 2. a-11
        Scene c(s); c.hold(0); c.hold(1);
        rs_frame_set_outliers(c.w, 0, s.outlier0.data());
        const int n = Matcher(0.9f, true).SearchByProjection(*c.F[1], *c.F[0], th[1], mono != 0);
        wr1(n); wr(c.ids(c.F[1]->mvpMapPoints));
    }
    {   // 3. a-12 (Cur, KF)
        Scene c(s); c.hold(0); c.hold(1);
        std::set<MapPoint*> found;
        for (int i = 0; i < nmp; ++i) if (s.found[i]) found.insert(c.mp[i]);
        const int n = Matcher(0.9f, true).SearchByProjection(*c.F[1], c.K[0], found, th[2], 64);
        wr1(n); wr(c.ids(c.F[1]->mvpMapPoints));
    }
    {   // 4. a-12 (KF, Scw)
        Scene c(s); c.hold(1);
        std::vector<MapPoint*> matched = c.K[1]->GetMapPointMatches();
        const int n = Matcher(0.75f, true).SearchByProjection(c.K[1], matf(4, 4, s.Scw.data()), c.mp, matched, (int)th[3]);
        wr1(n); wr(c.ids(matched));
    }
    {   // 5. a-13 (KF, F)
        Scene c(s); c.hold(0);
        std::vector<MapPoint*> bow;
        const int n = Matcher(0.75f, true).SearchByBoW(c.K[0], *c.F[1], bow);
        wr1(n); wr(c.ids(bow));
    }
    for (int only = 0; only < 2; ++only) {   // 6. a-14
        Scene c(s); c.hold(0); c.hold(1);
        std::vector<std::pair<size_t, size_t> > pairs;
        const int n = Matcher(0.6f, false).SearchForTriangulation(c.K[0], c.K[1], matf(3, 3, s.F12.data()), pairs, only != 0);
        std::vector<int> flat;
        for (size_t i = 0; i < pairs.size(); ++i) { flat.push_back((int)pairs[i].first); flat.push_back((int)pairs[i].second); }
        wr1(n); wr(flat);
    }
    {   // 7. Fuse(KF, vpMapPoints, th): the graph state the reference's MapPoint::Replace / AddObservation leave behind
        Scene c(s); c.hold(1, true);
        const int n = Matcher(0.8f, true).Fuse(c.K[1], c.mp, th[4]);
        wr1(n); wr(c.ids(c.K[1]->GetMapPointMatches()));
        std::vector<int> bad(nmp), nobs(nmp), repl(nmp);
        for (int i = 0; i < nmp; ++i) { bad[i] = c.mp[i]->isBad(); nobs[i] = c.mp[i]->Observations(); MapPoint* r = c.mp[i]->GetReplaced(); repl[i] = r ? rs_mp_id(c.w, r) : -1; }
        wr(bad); wr(nobs); wr(repl);
    }
    {   // 8. Fuse(KF, Scw, vpPoints, th, vpReplacePoint)
        Scene c(s); c.hold(1, true);
        std::vector<MapPoint*> repl(nmp, (MapPoint*)NULL);
        const int n = Matcher(0.8f, true).Fuse(c.K[1], matf(4, 4, s.Scw.data()), c.mp, th[5], repl);
        wr1(n); wr(c.ids(repl)); wr(c.ids(c.K[1]->GetMapPointMatches()));
    }
    {   // 9. SearchBySim3
        Scene c(s); c.hold(0); c.hold(1, true);
        std::vector<MapPoint*> m12(c.K[0]->N, (MapPoint*)NULL);
        for (int i = 0; i < c.K[0]->N; i += 17) {
            MapPoint* p = c.K[0]->GetMapPoint(i);
            if (p && p->IsInKeyFrame(c.K[1])) m12[i] = p;
        }
        const float s12 = s.sim[0];
        const int n = Matcher(0.8f, true).SearchBySim3(c.K[0], c.K[1], m12, s12, matf(3, 3, &s.sim[1]), matf(3, 1, &s.sim[10]), th[6]);
        wr1(n); wr(c.ids(m12));
    }
    {   // 10. SearchByBoW(KF1, KF2) - the MapFusion matcher
        Scene c(s); c.hold(0); c.hold(1);
        std::vector<MapPoint*> m12;
        const int n = Matcher(0.75f, true).SearchByBoW(c.K[0], c.K[1], m12);
        wr1(n); wr(c.ids(m12));
    }
    {   // 11. SearchForInitialization
        Scene c(s);
        std::vector<cv::Point2f> prev(c.F[0]->N);
        for (int i = 0; i < c.F[0]->N; ++i) prev[i] = c.F[0]->mvKeysUn[i].pt;
        std::vector<int> m12;
        const int n = Matcher(0.9f, true).SearchForInitialization(*c.F[0], *c.F[1], prev, m12, 100);
        wr1(n); wr(m12);
    }
    fclose(g_out);
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: real_types_test scenario.bin out_reference.bin out_b200.bin\n"); return 2; }
    const Scenario s = load(argv[1]);
    run_all<ORB_SLAM2::ORBmatcher>(s, argv[2]);
    if (std::string(argv[3]) != "-") run_all<ORB_SLAM2::B200ORBmatcher>(s, argv[3]);   // "-": reference pass only (no device)
    printf("real types ok\n");
    return 0;
}
