"""TEST INFRASTRUCTURE ONLY (oracle). ctypes binding of oracle/_ref/libref_slam.so: the reference's own ORBmatcher, Frame,
KeyFrame, MapPoint, Map and KeyFrameDatabase (compiled unmodified from /root/reference by oracle/build_ref.sh) driven from
arrays (oracle/ref_slam_wrap.cc). Used to pin oracle_lib.py's restatements and to generate tests/golden; never imported by
the product."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libref_slam.so")
f32 = np.float32
_L = None


def available():
    return os.path.exists(SO)


def L():
    global _L
    if _L is None:
        _L = C.CDLL(SO)
        _L.rs_create.restype = C.c_void_p
        for name in ("rs_norm_l2", "rs_dot"):
            getattr(_L, name).restype = C.c_double
    return _L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a, shape=None):
    a = np.ascontiguousarray(a, f32)
    return a if shape is None else a.reshape(shape)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


class World:
    """One Map + KeyFrameDatabase + vocabulary with index-addressed frames, keyframes and map points."""

    def __init__(self, voc_text=None):
        self.L = L()
        self.h = C.c_void_p(self.L.rs_create((voc_text or "").encode()))

    def __del__(self):
        try:
            self.L.rs_destroy(self.h)
        except Exception:
            pass

    # ---- frames ----
    def frame_arrays(self, kps6, desc, K4, width, height, uright=None, depth=None, dist=(0, 0, 0, 0), bf=40.0, th_depth=35.0,
                     scale_factor=1.2, nlevels=8, Tcw=None, sys=0):
        kps6 = _f(kps6); desc = np.ascontiguousarray(desc, np.uint8)
        ur = None if uright is None else _f(uright); dp = None if depth is None else _f(depth)
        T = None if Tcw is None else _f(Tcw, (4, 4))
        d = _f(dist)
        return self.L.rs_frame_arrays(self.h, sys, len(kps6), _p(kps6), _p(desc), _p(ur), _p(dp), _p(_f(K4)), _p(d), len(d), C.c_float(bf),
                                      C.c_float(th_depth), width, height, C.c_float(scale_factor), nlevels, _p(T))

    def frame_images(self, kind, img, K4, imgR=None, depth=None, dist=(0, 0, 0, 0), bf=40.0, th_depth=35.0, nfeatures=1000,
                     scale_factor=1.2, nlevels=8, ini=20, mn=7, sys=0, mb_before=0.0):
        img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
        r = None if imgR is None else np.ascontiguousarray(imgR, np.uint8)
        dm = None if depth is None else _f(depth)
        d = _f(dist)
        return self.L.rs_frame_images(self.h, sys, kind, _p(img), _p(r), _p(dm), w, h, _p(_f(K4)), _p(d), len(d), C.c_float(bf),
                                      C.c_float(th_depth), nfeatures, C.c_float(scale_factor), nlevels, ini, mn, C.c_float(mb_before))

    def frame_n(self, f):
        return self.L.rs_frame_n(self.h, f)

    def frame_get(self, f):
        n = self.frame_n(f)
        k = np.zeros((n, 6), f32); un = np.zeros((n, 2), f32); d = np.zeros((n, 32), np.uint8); ur = np.zeros(n, f32); dp = np.zeros(n, f32)
        self.L.rs_frame_get(self.h, f, _p(k), _p(un), _p(d), _p(ur), _p(dp))
        return dict(kps=k, un=un, desc=d, uright=ur, depth=dp)

    def frame_bounds(self, f=0):
        o = np.zeros(6, f32)
        self.L.rs_frame_bounds(self.h, f, _p(o))
        return o

    def frame_set_pose(self, f, Tcw):
        self.L.rs_frame_set_pose(self.h, f, _p(_f(Tcw, (4, 4))))

    def frame_set_mappoints(self, f, mp):
        self.L.rs_frame_set_mappoints(self.h, f, _p(_i(mp)))

    def frame_set_outliers(self, f, flags):
        self.L.rs_frame_set_outliers(self.h, f, _p(np.ascontiguousarray(flags, np.uint8)))

    def frame_get_mappoints(self, f):
        o = np.zeros(self.frame_n(f), np.int32)
        self.L.rs_frame_get_mappoints(self.h, f, _p(o))
        return o

    def frame_grid(self, f):
        counts = np.zeros(64 * 48, np.int32); items = np.zeros(max(1, self.frame_n(f)), np.int32)
        n = self.L.rs_frame_grid(self.h, f, _p(counts), _p(items))
        return counts.reshape(64, 48), items[:n]

    def frame_features_in_area(self, f, x, y, r, min_level=-1, max_level=-1):
        o = np.zeros(max(1, self.frame_n(f)), np.int32)
        n = self.L.rs_frame_features_in_area(self.h, f, C.c_float(x), C.c_float(y), C.c_float(r), min_level, max_level, _p(o), len(o))
        return o[:n].tolist()

    def kf_features_in_area(self, k, x, y, r, cap=8192):
        o = np.zeros(cap, np.int32)
        n = self.L.rs_kf_features_in_area(self.h, k, C.c_float(x), C.c_float(y), C.c_float(r), _p(o), cap)
        return o[:n].tolist()

    # ---- bag of words ----
    def _bow(self, fn, idx, cap):
        wi = np.zeros(cap, np.int32); ww = np.zeros(cap, np.float64); ni = np.zeros(cap, np.int32); no = np.zeros(cap + 1, np.int32)
        ft = np.zeros(cap, np.int32); nn = C.c_int(0)
        nw = fn(self.h, idx, _p(wi), _p(ww), _p(ni), _p(no), _p(ft), C.byref(nn))
        nn = nn.value
        fv = [(int(ni[i]), ft[no[i]:no[i + 1]].tolist()) for i in range(nn)]
        return wi[:nw].copy(), ww[:nw].copy(), fv

    def frame_compute_bow(self, f):
        return self._bow(self.L.rs_frame_compute_bow, f, self.frame_n(f) + 1)

    def kf_compute_bow(self, k, cap):
        return self._bow(self.L.rs_kf_compute_bow, k, cap + 1)

    @staticmethod
    def _fv_arrays(fv):
        ids = _i([n for n, _ in fv]); off = np.zeros(len(fv) + 1, np.int32)
        for i, (_, v) in enumerate(fv):
            off[i + 1] = off[i] + len(v)
        feat = _i([x for _, v in fv for x in v]) if len(fv) else np.zeros(1, np.int32)
        return ids, off, feat

    def frame_set_featvec(self, f, fv):
        ids, off, feat = self._fv_arrays(fv)
        self.L.rs_frame_set_featvec(self.h, f, len(fv), _p(ids), _p(off), _p(feat))

    def kf_set_featvec(self, k, fv):
        ids, off, feat = self._fv_arrays(fv)
        self.L.rs_kf_set_featvec(self.h, k, len(fv), _p(ids), _p(off), _p(feat))

    def frame_set_bowvec(self, f, ids, wt):
        ids = _i(ids); wt = np.ascontiguousarray(wt, np.float64)
        self.L.rs_frame_set_bowvec(self.h, f, len(ids), _p(ids), _p(wt))

    def kf_set_scores(self, k, covis=0.0, reloc=0.0):
        self.L.rs_kf_set_scores(self.h, k, C.c_float(covis), C.c_float(reloc))

    def kf_reloc_score(self, k):
        self.L.rs_kf_reloc_score.restype = C.c_float
        return float(self.L.rs_kf_reloc_score(self.h, k))

    def kf_set_bowvec(self, k, ids, wt):
        ids = _i(ids); wt = np.ascontiguousarray(wt, np.float64)
        self.L.rs_kf_set_bowvec(self.h, k, len(ids), _p(ids), _p(wt))

    # ---- keyframes / map points ----
    def keyframe(self, f, sys=0):
        return self.L.rs_keyframe(self.h, f, sys)

    def kf_set_pose(self, k, Tcw):
        self.L.rs_kf_set_pose(self.h, k, _p(_f(Tcw, (4, 4))))

    def kf_get_mappoints(self, k, n):
        o = np.zeros(n, np.int32)
        self.L.rs_kf_get_mappoints(self.h, k, _p(o))
        return o

    def kf_set_mappoints(self, k, mp):
        self.L.rs_kf_set_mappoints(self.h, k, _p(_i(mp)))

    def kf_add_connection(self, a, b, weight):
        self.L.rs_kf_add_connection(self.h, a, b, weight)

    def mappoint(self, pos, ref_kf):
        return self.L.rs_mappoint(self.h, _p(_f(pos)), ref_kf)

    def mappoint_from_frame(self, pos, f, idx):
        return self.L.rs_mappoint_from_frame(self.h, _p(_f(pos)), f, idx)

    def mp_set(self, m, desc=None, normal=None, min_max=None, n_obs=None, bad=None):
        d = None if desc is None else np.ascontiguousarray(desc, np.uint8)
        nr = None if normal is None else _f(normal); mm = None if min_max is None else _f(min_max)
        no = None if n_obs is None else C.byref(C.c_int(int(n_obs))); bd = None if bad is None else C.byref(C.c_int(int(bad)))
        self.L.rs_mp_set(self.h, m, _p(d), _p(nr), _p(mm), no, bd)

    def mp_get(self, m):
        pos = np.zeros(3, f32); d = np.zeros(32, np.uint8); nr = np.zeros(3, f32); mm = np.zeros(2, f32); no = C.c_int(0); bd = C.c_int(0)
        self.L.rs_mp_get(self.h, m, _p(pos), _p(d), _p(nr), _p(mm), C.byref(no), C.byref(bd))
        return dict(pos=pos, desc=d, normal=nr, min_max=mm, n_obs=no.value, bad=bool(bd.value))

    def observe(self, m, k, idx):
        self.L.rs_observe(self.h, m, k, idx)

    def mp_set_observation(self, m, k, idx):
        self.L.rs_mp_set_observation(self.h, m, k, int(idx))

    def mp_compute_distinctive(self, m):
        self.L.rs_mp_compute_distinctive(self.h, m)

    def mp_update_normal_and_depth(self, m):
        self.L.rs_mp_update_normal_and_depth(self.h, m)

    def mp_set_track(self, m, in_view, px, py, pxr, level, view_cos, sys=0):
        self.L.rs_mp_set_track(self.h, m, sys, int(in_view), C.c_float(px), C.c_float(py), C.c_float(pxr), int(level), C.c_float(view_cos))

    def is_in_frustum(self, f, m, cos_limit=0.5):
        o = np.zeros(6, f32)
        self.L.rs_is_in_frustum(self.h, f, m, C.c_float(cos_limit), _p(o))
        return o

    def mp_replaced(self, m):
        return self.L.rs_mp_replaced(self.h, m)

    def mp_index_in_kf(self, m, k):
        return self.L.rs_mp_index_in_kf(self.h, m, k)

    # ---- searches ----
    def search_by_projection_local(self, f, mp_ids, th, nnratio=0.8):
        ids = _i(mp_ids)
        return self.L.rs_search_by_projection_local(self.h, C.c_float(nnratio), f, _p(ids), len(ids), C.c_float(th))

    def search_by_projection_last(self, cur, last, th, mono, nnratio=0.9, check_ori=True):
        return self.L.rs_search_by_projection_last(self.h, C.c_float(nnratio), int(check_ori), cur, last, C.c_float(th), int(mono))

    def search_by_projection_kf(self, cur, kf, found, th, orb_dist, nnratio=0.9, check_ori=True):
        fd = _i(found) if len(found) else np.zeros(1, np.int32)
        return self.L.rs_search_by_projection_kf(self.h, C.c_float(nnratio), int(check_ori), cur, kf, _p(fd), len(found), C.c_float(th), int(orb_dist))

    def search_by_projection_sim3(self, kf, Scw, pts, matched, th, nnratio=0.75):
        pts = _i(pts); m = _i(matched).copy()
        r = self.L.rs_search_by_projection_sim3(self.h, C.c_float(nnratio), kf, _p(_f(Scw, (4, 4))), _p(pts), len(pts), _p(m), int(th))
        return r, m

    def search_by_bow_kf_f(self, kf, f, nnratio=0.7, check_ori=True):
        o = np.zeros(self.frame_n(f), np.int32)
        r = self.L.rs_search_by_bow_kf_f(self.h, C.c_float(nnratio), int(check_ori), kf, f, _p(o))
        return r, o

    def search_by_bow_kf_kf(self, k1, k2, n1, nnratio=0.75, check_ori=True):
        o = np.zeros(n1, np.int32)
        r = self.L.rs_search_by_bow_kf_kf(self.h, C.c_float(nnratio), int(check_ori), k1, k2, _p(o))
        return r, o

    def search_for_initialization(self, f1, f2, prev_matched, window, nnratio=0.9, check_ori=True):
        pm = _f(prev_matched).copy(); m = np.zeros(self.frame_n(f1), np.int32)
        r = self.L.rs_search_for_initialization(self.h, C.c_float(nnratio), int(check_ori), f1, f2, _p(pm), _p(m), int(window))
        return r, m, pm

    def search_for_triangulation(self, k1, k2, F12, only_stereo, nnratio=0.6, check_ori=True, cap=8192):
        o = np.zeros((cap, 2), np.int32)
        n = self.L.rs_search_for_triangulation(self.h, C.c_float(nnratio), int(check_ori), k1, k2, _p(_f(F12, (3, 3))), int(only_stereo), _p(o), cap)
        return [tuple(int(x) for x in r) for r in o[:n]]

    def search_by_sim3(self, k1, k2, matches12, s12, R12, t12, th, nnratio=0.75):
        m = _i(matches12).copy()
        r = self.L.rs_search_by_sim3(self.h, C.c_float(nnratio), k1, k2, _p(m), C.c_float(s12), _p(_f(R12, (3, 3))), _p(_f(t12)), C.c_float(th))
        return r, m

    def fuse(self, kf, pts, th, nnratio=0.6):
        pts = _i(pts)
        return self.L.rs_fuse(self.h, C.c_float(nnratio), kf, _p(pts), len(pts), C.c_float(th))

    def fuse_sim3(self, kf, Scw, pts, th, nnratio=0.8):
        pts = _i(pts); rp = np.full(len(pts), -1, np.int32)
        r = self.L.rs_fuse_sim3(self.h, C.c_float(nnratio), kf, _p(_f(Scw, (4, 4))), _p(pts), len(pts), C.c_float(th), _p(rp))
        return r, rp

    # ---- keyframe database ----
    def db_add(self, k):
        self.L.rs_db_add(self.h, k)

    def db_erase(self, k):
        self.L.rs_db_erase(self.h, k)

    def db_detect_loop_candidates(self, k, min_score, cap=4096):
        o = np.zeros(cap, np.int32)
        n = self.L.rs_db_detect_loop_candidates(self.h, k, C.c_float(min_score), _p(o), cap)
        return o[:n].tolist()

    def db_detect_relocalization_candidates(self, f, cap=4096):
        o = np.zeros(cap, np.int32)
        n = self.L.rs_db_detect_relocalization_candidates(self.h, f, _p(o), cap)
        return o[:n].tolist()

    def db_detect_covisibility_candidates(self, k, min_score, ignore=(), cap=4096):
        o = np.zeros(cap, np.int32); ig = _i(list(ignore)) if len(ignore) else np.zeros(1, np.int32)
        n = self.L.rs_db_detect_covisibility_candidates(self.h, k, C.c_float(min_score), _p(ig), len(ignore), _p(o), cap)
        return o[:n].tolist()


# ---- the stand-in's float cv::Mat arithmetic, for pinning against cv2 ----
def gemm32f(A, B, alpha=1.0, Cm=None, beta=0.0, flags=0):
    A = _f(A); B = _f(B)
    m = A.shape[1] if flags & 1 else A.shape[0]; n = B.shape[0] if flags & 2 else B.shape[1]
    D = np.zeros((m, n), f32)
    Cc = None if Cm is None else _f(Cm)
    L().rs_gemm32f(_p(A), A.shape[0], A.shape[1], _p(B), B.shape[0], B.shape[1], C.c_double(alpha), _p(Cc), C.c_double(beta), _p(D), flags)
    return D


def norm_l2(v):
    v = _f(v).ravel()
    return L().rs_norm_l2(_p(v), len(v))


def dot(a, b):
    a = _f(a).ravel(); b = _f(b).ravel()
    return L().rs_dot(_p(a), _p(b), len(a))


def undistort_points(pts, K, dist):
    pts = _f(pts, (-1, 2)); out = np.zeros_like(pts); K = _f(K, (3, 3)); d = _f(dist)
    L().rs_undistort_points(_p(pts), _p(out), len(pts), _p(K), _p(d), len(d))
    return out
