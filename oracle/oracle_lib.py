"""ctypes bindings of the CPU oracle (oracle/liborb_oracle.so) and of the reference extractor
compiled against the OpenCV stand-in (oracle/_ref/*.so). Test infrastructure only."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in oracle/)
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def pattern():
    txt = open(os.path.join(ROOT, "multiagent_orb_slam2_b200", "csrc", "orb_pattern_31.inc")).read()
    txt = "\n".join(l for l in txt.splitlines() if not l.startswith("//"))
    v = np.array([int(x) for x in re.findall(r"-?\d+", txt)], np.int32)
    assert v.size == 1024
    return v


def lib():
    global _lib
    try:
        return _lib
    except NameError:
        pass
    L = C.CDLL(os.path.join(ROOT, "oracle", "liborb_oracle.so"))
    L.orc_fast_atan2.restype = C.c_float
    L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.orc_round.argtypes = [C.c_float]
    L.orc_extractor_create.restype = C.c_void_p
    L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.orc_extractor_destroy.argtypes = [C.c_void_p]
    L.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
    L.orc_get_result.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_get_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    L.orc_level_dims.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4
    L.orc_level_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.orc_level_has_blur.argtypes = [C.c_void_p, C.c_int]
    L.orc_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_level_selected.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_extract_mt.restype = C.c_long
    L.orc_stereo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def resize(src, dw, dh):
    src = np.ascontiguousarray(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(src.ctypes.data_as(C.c_void_p), src.shape[1], src.shape[0], C.c_size_t(src.strides[0]),
                               dst.ctypes.data_as(C.c_void_p), dw, dh, C.c_size_t(dw))
    return dst


def gaussian(src):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_gaussian7x7_u8(src.ctypes.data_as(C.c_void_p), src.shape[1], src.shape[0], C.c_size_t(src.strides[0]),
                             dst.ctypes.data_as(C.c_void_p), C.c_size_t(dst.strides[0]))
    return dst


def border(src, b=19):
    src = np.ascontiguousarray(src)
    dst = np.empty((src.shape[0] + 2 * b, src.shape[1] + 2 * b), np.uint8)
    lib().orc_copy_make_border(src.ctypes.data_as(C.c_void_p), src.shape[1], src.shape[0], C.c_size_t(src.strides[0]),
                               dst.ctypes.data_as(C.c_void_p), C.c_size_t(dst.strides[0]), b)
    return dst


def fast(img, th):
    """img may be a (non-contiguous) view with unit column stride."""
    assert img.strides[1] == 1
    cap = img.size
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_fast9_nms(C.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]), th,
                            out.ctypes.data_as(C.c_void_p), cap)
    return out[:n]


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def quadtree(xs, ys, scores, minX, maxX, minY, maxY, N):
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    sc = np.ascontiguousarray(scores, np.int32)
    out = np.empty(max(len(xs), 1), np.int32)
    n = lib().orc_quadtree(xs.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                           len(xs), minX, maxX, minY, maxY, N, out.ctypes.data_as(C.c_void_p), len(out))
    return out[:n]


class OracleExtractor:
    """Mirror of ORB_SLAM2::ORBextractor over the oracle, with access to every intermediate."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7):
        self.L = lib()
        self.pat = pattern()
        self.nlevels = nlevels
        self.h = self.L.orc_extractor_create(nfeatures, scale, nlevels, ini, mn, self.pat.ctypes.data_as(C.c_void_p))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_extractor_destroy(self.h)
            self.h = None

    def __call__(self, img):
        img = np.ascontiguousarray(img)
        n = self.L.orc_extract(self.h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]))
        kps = np.empty((n, 6), np.float32)
        desc = np.empty((n, 32), np.uint8)
        self.L.orc_get_result(self.h, kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p))
        return kps, desc

    def tables(self):
        n = self.nlevels
        t = [np.empty(n, np.float32) for _ in range(4)] + [np.empty(n, np.int32), np.empty(16, np.int32)]
        self.L.orc_get_tables(self.h, *[a.ctypes.data_as(C.c_void_p) for a in t])
        return dict(zip(["scale", "inv_scale", "sigma2", "inv_sigma2", "quota", "umax"], t))

    def level(self, l):
        w, h, nc, ns = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.L.orc_level_dims(self.h, l, C.byref(w), C.byref(h), C.byref(nc), C.byref(ns))
        img = np.empty((h.value, w.value), np.uint8)
        self.L.orc_level_image(self.h, l, img.ctypes.data_as(C.c_void_p), 0)
        blur = None
        if self.L.orc_level_has_blur(self.h, l):
            blur = np.empty_like(img)
            self.L.orc_level_image(self.h, l, blur.ctypes.data_as(C.c_void_p), 1)
        cand = np.empty((nc.value, 3), np.int32)
        self.L.orc_level_candidates(self.h, l, cand.ctypes.data_as(C.c_void_p))
        sel = np.empty(ns.value, np.int32)
        ang = np.empty(ns.value, np.float32)
        self.L.orc_level_selected(self.h, l, sel.ctypes.data_as(C.c_void_p), ang.ctypes.data_as(C.c_void_p))
        return dict(img=img, blur=blur, cand=cand, sel=sel, angle=ang)


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_hamming(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))


def knn2(A, B, threads=1):
    A = np.ascontiguousarray(A, np.uint8); B = np.ascontiguousarray(B, np.uint8)
    idx = np.empty(len(A), np.int32); d1 = np.empty(len(A), np.int32); d2 = np.empty(len(A), np.int32)
    lib().orc_knn2_mt(A.ctypes.data_as(C.c_void_p), len(A), B.ctypes.data_as(C.c_void_p), len(B),
                      idx.ctypes.data_as(C.c_void_p), d1.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p), threads)
    return idx, d1, d2


def knn2_lists(A, B, offsets, cands):
    A = np.ascontiguousarray(A, np.uint8); B = np.ascontiguousarray(B, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32); cands = np.ascontiguousarray(cands, np.int32)
    idx = np.empty(len(A), np.int32); d1 = np.empty(len(A), np.int32); d2 = np.empty(len(A), np.int32)
    lib().orc_knn2_lists(A.ctypes.data_as(C.c_void_p), len(A), B.ctypes.data_as(C.c_void_p),
                         offsets.ctypes.data_as(C.c_void_p), cands.ctypes.data_as(C.c_void_p),
                         idx.ctypes.data_as(C.c_void_p), d1.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p))
    return idx, d1, d2


# ---- the reference's own extractor (oracle/_ref), when it was built ------------------------
def ref_available(kind="canonical"):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_orb_%s.so" % kind))


class RefExtractor:
    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, kind="canonical"):
        self.L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_orb_%s.so" % kind))
        self.L.ref_extractor_create.restype = C.c_void_p
        self.L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        self.L.ref_extractor_destroy.argtypes = [C.c_void_p]
        self.L.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        self.L.ref_get_tables.argtypes = [C.c_void_p] * 5
        self.L.ref_level_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.L.ref_level_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.nlevels = nlevels
        self.cap = nfeatures + 4 * nlevels + 64
        self.h = self.L.ref_extractor_create(nfeatures, scale, nlevels, ini, mn)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_extractor_destroy(self.h)
            self.h = None

    def __call__(self, img):
        img = np.ascontiguousarray(img)
        kps = np.empty((self.cap, 6), np.float32)
        desc = np.empty((self.cap, 32), np.uint8)
        n = self.L.ref_extract(self.h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]),
                               kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), self.cap)
        assert n <= self.cap
        return kps[:n].copy(), desc[:n].copy()

    def tables(self):
        t = [np.empty(self.nlevels, np.float32) for _ in range(4)]
        self.L.ref_get_tables(self.h, *[a.ctypes.data_as(C.c_void_p) for a in t])
        return dict(zip(["scale", "inv_scale", "sigma2", "inv_sigma2"], t))

    def level_image(self, l, border=0):
        w, h = C.c_int(), C.c_int()
        self.L.ref_level_dims(self.h, l, C.byref(w), C.byref(h))
        out = np.empty((h.value + 2 * border, w.value + 2 * border), np.uint8)
        self.L.ref_level_image(self.h, l, out.ctypes.data_as(C.c_void_p), border)
        return out


def quadtree_arrayform(xs, ys, scores, minX, maxX, minY, maxY, N):
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    sc = np.ascontiguousarray(scores, np.int32)
    out = np.empty(max(len(xs), 1), np.int32)
    n = lib().orc_quadtree_arrayform(xs.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p),
                                     sc.ctypes.data_as(C.c_void_p), len(xs), minX, maxX, minY, maxY, N,
                                     out.ctypes.data_as(C.c_void_p), len(out))
    return out[:n]


# ---- matcher restatements (pure Python loops: small cases only) --------------------------------
class OracleFrame:
    """The Frame members the windowed searches touch: undistorted keypoints, descriptors and the
    64 x 48 grid (AssignFeaturesToGrid src/Frame.cc:230-245, GetFeaturesInArea 327-380)."""
    COLS, ROWS = 64, 48

    def __init__(self, kps6, desc, width, height, bounds=None):
        """kps6: the UNDISTORTED keypoints (mvKeysUn); bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY) of image_bounds() when the
        camera has distortion, else the image rectangle (src/Frame.cc:436-464)."""
        self.kps, self.desc = kps6, desc
        b = (0, width, 0, height) if bounds is None else bounds
        self.minX, self.minY = np.float32(b[0]), np.float32(b[2])
        self.wInv = np.float32(self.COLS) / np.float32(np.float32(b[1]) - np.float32(b[0]))
        self.hInv = np.float32(self.ROWS) / np.float32(np.float32(b[3]) - np.float32(b[2]))
        self.grid = [[[] for _ in range(self.ROWS)] for _ in range(self.COLS)]
        for i in range(len(kps6)):
            # C round(): half away from zero
            fx = float(np.float32(np.float32(kps6[i, 0] - self.minX) * self.wInv)); fy = float(np.float32(np.float32(kps6[i, 1] - self.minY) * self.hInv))
            px = int(np.floor(fx + 0.5)) if fx >= 0 else -int(np.floor(-fx + 0.5))
            py = int(np.floor(fy + 0.5)) if fy >= 0 else -int(np.floor(-fy + 0.5))
            if 0 <= px < self.COLS and 0 <= py < self.ROWS:
                self.grid[px][py].append(i)

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        x, y, r = np.float32(x), np.float32(y), np.float32(r)
        out = []
        c0 = max(0, int(np.floor(np.float32(np.float32(x - self.minX - r) * self.wInv))))
        if c0 >= self.COLS:
            return out
        c1 = min(self.COLS - 1, int(np.ceil(np.float32(np.float32(x - self.minX + r) * self.wInv))))
        if c1 < 0:
            return out
        r0 = max(0, int(np.floor(np.float32(np.float32(y - self.minY - r) * self.hInv))))
        if r0 >= self.ROWS:
            return out
        r1 = min(self.ROWS - 1, int(np.ceil(np.float32(np.float32(y - self.minY + r) * self.hInv))))
        if r1 < 0:
            return out
        check = min_level > 0 or max_level >= 0
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for j in self.grid[ix][iy]:
                    o = int(self.kps[j, 5])
                    if check and (o < min_level or (max_level >= 0 and o > max_level)):
                        continue
                    if abs(np.float32(self.kps[j, 0] - x)) < r and abs(np.float32(self.kps[j, 1] - y)) < r:
                        out.append(j)
        return out


def undistort_points(pts, K, dist):
    """cv::undistortPoints(pts, pts, K, dist, noArray(), K) as Frame::UndistortKeyPoints calls it (src/Frame.cc:404-434):
    5 fixed-point iterations of the distortion model in double, re-projection with K, result stored as float
    (pinned bit-exactly to cv2 4.13 in tests/test_oracle_vs_cv2.py)."""
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    K = np.asarray(K, np.float32).reshape(3, 3).astype(np.float64)
    k = np.zeros(12); d = np.asarray(dist, np.float32).astype(np.float64).ravel(); k[:len(d)] = d
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ifx, ify = 1.0 / fx, 1.0 / fy
    out = np.empty_like(pts)
    for i in range(len(pts)):
        x0 = (float(pts[i, 0]) - cx) * ifx; y0 = (float(pts[i, 1]) - cy) * ify
        x, y = x0, y0
        for _ in range(5):
            r2 = x * x + y * y
            icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            if icdist < 0:
                x, y = x0, y0
                break
            dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
            dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
            x = (x0 - dx) * icdist; y = (y0 - dy) * icdist
        out[i, 0] = np.float32(fx * x + cx); out[i, 1] = np.float32(fy * y + cy)
    return out


def image_bounds(width, height, K, dist):
    """Frame::ComputeImageBounds, src/Frame.cc:436-464: (mnMinX, mnMaxX, mnMinY, mnMaxY) as floats."""
    f = np.float32
    if np.float32(np.asarray(dist).ravel()[0]) == 0:
        return np.array([0, width, 0, height], f)
    c = undistort_points(np.array([[0, 0], [width, 0], [0, height], [width, height]], f), K, dist)
    return np.array([min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])], f)


def stereo_from_rgbd(kps, un, depth, mbf):
    """Frame::ComputeStereoFromRGBD, src/Frame.cc:643-664: depth sampled at the truncated RAW keypoint position, the right
    coordinate from the UNDISTORTED x."""
    f = np.float32
    n = len(kps)
    ur, dz = np.full(n, -1, f), np.full(n, -1, f)
    for i in range(n):
        d = depth[int(kps[i, 1]), int(kps[i, 0])]
        if d > 0:
            dz[i] = d
            ur[i] = f(f(un[i, 0]) - f(f(mbf) / f(d)))
    return ur, dz


def three_maxima(sizes):
    """ComputeThreeMaxima, src/ORBmatcher.cc:1603-1644."""
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > m1:
            m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2:
            m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3:
            m3, i3 = s, i
    lim = np.float32(0.1) * np.float32(m1)
    if np.float32(m2) < lim:
        i2 = i3 = -1
    elif np.float32(m3) < lim:
        i3 = -1
    return i1, i2, i3


def search_for_initialization(F1, F2, prev_matched, window, nnratio=0.9, check_ori=True, th_low=50, histo=30):
    """ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:407-522. prev_matched: (n1,2) float32,
    updated in place. Returns (nmatches, vnMatches12)."""
    INT_MAX = 2**31 - 1
    n1, n2 = len(F1.kps), len(F2.kps)
    m12 = np.full(n1, -1, np.int32)
    m21 = np.full(n2, -1, np.int32)
    mdist = np.full(n2, INT_MAX, np.int64)
    rot = [[] for _ in range(histo)]
    factor = np.float32(1.0) / np.float32(histo)
    nm = 0
    ratio = np.float32(nnratio)
    for i1 in range(n1):
        if int(F1.kps[i1, 5]) > 0:
            continue
        cand = F2.features_in_area(prev_matched[i1, 0], prev_matched[i1, 1], window, 0, 0)
        if not cand:
            continue
        best = best2 = INT_MAX
        bidx = -1
        for i2 in cand:
            d = hamming(F1.desc[i1], F2.desc[i2])
            if mdist[i2] <= d:
                continue
            if d < best:
                best2, best, bidx = best, d, i2
            elif d < best2:
                best2 = d
        if best <= th_low:
            if np.float32(best) < np.float32(best2) * ratio:
                if m21[bidx] >= 0:
                    m12[m21[bidx]] = -1
                    nm -= 1
                m12[i1] = bidx
                m21[bidx] = i1
                mdist[bidx] = best
                nm += 1
                if check_ori:
                    r = np.float32(F1.kps[i1, 3] - F2.kps[bidx, 3])
                    if r < 0:
                        r = np.float32(r + np.float32(360.0))
                    v = float(np.float32(r * factor))
                    b = int(np.floor(v + 0.5))
                    if b == histo:
                        b = 0
                    rot[b].append(i1)
    if check_ori:
        keep = three_maxima([len(x) for x in rot])
        for i in range(histo):
            if i in keep:
                continue
            for idx1 in rot[i]:
                if m12[idx1] >= 0:
                    m12[idx1] = -1
                    nm -= 1
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev_matched[i1] = F2.kps[m12[i1], :2]
    return nm, m12


def stereo_match(oL, oR, mbf, mb):
    """Frame::ComputeStereoMatches on the last results of two OracleExtractor objects (left, right)."""
    n = lib().orc_extract  # noqa: F841 (bindings loaded)
    w, h, nc, ns = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    nL = 0
    for l in range(oL.nlevels):
        lib().orc_level_dims(oL.h, l, C.byref(w), C.byref(h), C.byref(nc), C.byref(ns))
        nL += ns.value
    u = np.empty(max(nL, 1), np.float32); d = np.empty(max(nL, 1), np.float32)
    kept = lib().orc_stereo_match(oL.h, oR.h, mbf, mb, u.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p))
    return u[:nL], d[:nL], kept


def search_by_bow_kf_kf(desc1, fv1, valid1, ang1, desc2, fv2, valid2, ang2, nnratio=0.75, check_ori=True, th_low=50, histo=30):
    """ORBmatcher::SearchByBoW(KF1, KF2), src/ORBmatcher.cc:524-657; feature vectors as sorted
    (node, [indices]) lists, valid* = map point present and not bad."""
    n1, n2 = len(desc1), len(desc2)
    m12 = np.full(n1, -1, np.int32)
    matched2 = np.zeros(n2, bool)
    rot = [[] for _ in range(histo)]
    factor = np.float32(1.0) / np.float32(histo)
    ratio = np.float32(nnratio)
    nm = 0
    i = j = 0
    while i < len(fv1) and j < len(fv2):
        if fv1[i][0] == fv2[j][0]:
            for idx1 in fv1[i][1]:
                if not valid1[idx1]:
                    continue
                b1 = b2 = 256
                bidx = -1
                for idx2 in fv2[j][1]:
                    if matched2[idx2] or not valid2[idx2]:
                        continue
                    d = hamming(desc1[idx1], desc2[idx2])
                    if d < b1:
                        b2, b1, bidx = b1, d, idx2
                    elif d < b2:
                        b2 = d
                if b1 < th_low and np.float32(b1) < ratio * np.float32(b2):
                    m12[idx1] = bidx
                    matched2[bidx] = True
                    if check_ori:
                        r = np.float32(np.float32(ang1[idx1]) - np.float32(ang2[bidx]))
                        if r < 0:
                            r = np.float32(r + np.float32(360.0))
                        b = int(np.floor(float(np.float32(r * factor)) + 0.5))
                        rot[0 if b == histo else b].append(idx1)
                    nm += 1
            i += 1; j += 1
        elif fv1[i][0] < fv2[j][0]:
            while i < len(fv1) and fv1[i][0] < fv2[j][0]:  # lower_bound
                i += 1
        else:
            while j < len(fv2) and fv2[j][0] < fv1[i][0]:
                j += 1
    if check_ori:
        keep = three_maxima([len(x) for x in rot])
        for b in range(histo):
            if b not in keep:
                for idx1 in rot[b]:
                    m12[idx1] = -1
                    nm -= 1
    return nm, m12


def distinctive_descriptor(descs):
    """MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:246-311."""
    n = len(descs)
    best, best_i = 2**31 - 1, 0
    for i in range(n):
        v = sorted(0 if i == j else hamming(descs[i], descs[j]) for j in range(n))
        med = v[int(0.5 * (n - 1))]
        if med < best:
            best, best_i = med, i
    return best_i


# ---- DBoW2 vocabulary transform (oracle restatement + the vendored DBoW2 itself) -----------------
class OracleVocabulary:
    def __init__(self, voc):
        L = lib()
        L.orc_voc_create.restype = C.c_void_p
        L.orc_voc_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_voc_destroy.argtypes = [C.c_void_p]
        L.orc_voc_descend.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        self.L = L
        parent = np.ascontiguousarray(voc["parent"], np.int32)
        desc = np.ascontiguousarray(voc["desc"], np.uint8)
        weight = np.ascontiguousarray(voc["weight"], np.float64)
        self.h = L.orc_voc_create(parent.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), weight.ctypes.data_as(C.c_void_p),
                                  len(parent), voc["k"], voc["L"])
        self.fn_descend, self.fn_transform = L.orc_voc_descend, L.orc_voc_transform

    def __del__(self):
        if getattr(self, "h", None) and hasattr(self.L, "orc_voc_destroy"):
            self.L.orc_voc_destroy(self.h)
            self.h = None

    def descend(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8)
        n = len(d)
        word = np.empty(n, np.int32); node = np.empty(n, np.int32); weight = np.empty(n, np.float64)
        self.fn_descend(self.h, d.ctypes.data_as(C.c_void_p), n, levelsup, word.ctypes.data_as(C.c_void_p),
                        node.ctypes.data_as(C.c_void_p), weight.ctypes.data_as(C.c_void_p))
        return word, node, weight

    def transform(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8)
        n = len(d)
        bid = np.empty(n + 1, np.int32); bval = np.empty(n + 1, np.float64)
        fnode = np.empty(n + 1, np.int32); foff = np.empty(n + 2, np.int32); ffeat = np.empty(n + 1, np.int32)
        nb, nf = C.c_int(), C.c_int()
        self.fn_transform(self.h, d.ctypes.data_as(C.c_void_p), n, levelsup, bid.ctypes.data_as(C.c_void_p), bval.ctypes.data_as(C.c_void_p),
                          n + 1, C.byref(nb), fnode.ctypes.data_as(C.c_void_p), foff.ctypes.data_as(C.c_void_p),
                          ffeat.ctypes.data_as(C.c_void_p), n + 1, C.byref(nf))
        fv = [(int(fnode[i]), ffeat[foff[i]:foff[i + 1]].tolist()) for i in range(nf.value)]
        return bid[:nb.value].copy(), bval[:nb.value].copy(), fv


class RefVocabulary(OracleVocabulary):
    """The reference's vendored DBoW2 (oracle/_ref/libref_dbow.so), loaded from an ORBvoc-format text file."""

    def __init__(self, text_file):
        L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_dbow.so"))
        L.refv_load.restype = C.c_void_p
        L.refv_load.argtypes = [C.c_char_p]
        L.refv_destroy.argtypes = [C.c_void_p]
        L.refv_descend.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refv_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        self.L = L
        self.h = L.refv_load(text_file.encode())
        assert self.h, "vocabulary text file rejected by DBoW2"
        self.fn_descend, self.fn_transform = L.refv_descend, L.refv_transform

    def size(self):
        return self.L.refv_size(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.refv_destroy(self.h)
            self.h = None


def dbow_ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_dbow.so"))


def search_by_projection_frame_mappoints(F, mp, th, nnratio=0.8, th_high=100):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th), src/ORBmatcher.cc:45-131, on arrays.
    F: OracleFrame plus F.scale (mvScaleFactors), F.uright (mvuRight), F.occupied (bool per keypoint: holds a map point
    with observations > 0). mp: dict of per-map-point arrays: in_view, bad, level (mnTrackScaleLevel), view_cos,
    proj_x, proj_y, proj_xr, desc. Returns (nmatches, assigned) with assigned[kp] = map point index or -1; keypoints
    assigned here count as occupied for the following map points (they get a map point with observations)."""
    n_kp = len(F.kps)
    assigned = np.full(n_kp, -1, np.int32)
    occupied = F.occupied.copy()
    nm = 0
    bfactor = th != 1.0
    ratio = np.float32(nnratio)
    for i in range(len(mp["level"])):
        if not mp["in_view"][i] or mp["bad"][i]:
            continue
        lvl = int(mp["level"][i])
        r = np.float32(2.5) if mp["view_cos"][i] > 0.998 else np.float32(4.0)
        if bfactor:
            r = np.float32(r * np.float32(th))
        rad = np.float32(r * F.scale[lvl])
        cand = F.features_in_area(mp["proj_x"][i], mp["proj_y"][i], rad, lvl - 1, lvl)
        if not cand:
            continue
        b1 = b2 = 256
        l1 = l2 = bi = -1
        for idx in cand:
            if occupied[idx]:
                continue
            if F.uright[idx] > 0:
                er = abs(np.float32(mp["proj_xr"][i]) - np.float32(F.uright[idx]))
                if er > rad:
                    continue
            d = hamming(mp["desc"][i], F.desc[idx])
            if d < b1:
                b2, b1, l2, l1, bi = b1, d, l1, int(F.kps[idx, 5]), idx
            elif d < b2:
                l2, b2 = int(F.kps[idx, 5]), d
        if b1 <= th_high:
            if l1 == l2 and np.float32(b1) > ratio * np.float32(b2):
                continue
            assigned[bi] = i
            occupied[bi] = bool(mp["has_obs"][i]) if "has_obs" in mp else True
            nm += 1
    return nm, assigned


# ---- the reference's own DescriptorDistance / ComputeThreeMaxima (oracle/_ref/libref_matcher_bits.so) ----
def matcher_bits_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_matcher_bits.so"))


class RefMatcherBits:
    """src/ORBmatcher.cc:1603-1665 compiled from the reference file itself (see build_ref.sh)."""

    def __init__(self):
        self.L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_matcher_bits.so"))
        self.L.refm_descriptor_distance.restype = C.c_int

    def descriptor_distance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(32); b = np.ascontiguousarray(b, np.uint8).reshape(32)
        return int(self.L.refm_descriptor_distance(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)))

    def three_maxima(self, sizes):
        s = np.ascontiguousarray(sizes, np.int32); out = np.zeros(3, np.int32)
        self.L.refm_three_maxima(s.ctypes.data_as(C.c_void_p), len(s), out.ctypes.data_as(C.c_void_p))
        return tuple(int(x) for x in out)


# ---- the remaining guided searches (rows a-11 ... a-15), scalar restatements on arrays ------------------
def _c_round(x):
    x = float(x)
    return int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5))


def _rot_bin(a1, a2, L=30):
    rot = np.float32(np.float32(a1) - np.float32(a2))
    if rot < 0.0:
        rot = np.float32(rot + np.float32(360.0))
    b = _c_round(np.float32(rot * (np.float32(1.0) / np.float32(L))))
    return 0 if b == L else b


def _rot_cleanup(rot, clear, L=30):
    keep = three_maxima([len(x) for x in rot])
    removed = 0
    for i in range(L):
        if i not in keep:
            for j in rot[i]:
                clear(j)
                removed += 1
    return removed


def search_by_projection_cur_last(F, last, th, forward, backward, check_ori=True, th_high=100, cleared=None):
    """src/ORBmatcher.cc:1330-1472. F: OracleFrame (current) with F.scale, F.uright, F.occupied; last: dict of arrays
    valid, u, v, ur, octave, angle, desc, has_obs."""
    assigned = np.full(len(F.kps), -1, np.int32)
    occupied = F.occupied.copy()
    rot = [[] for _ in range(30)]
    nm = 0
    for i in range(len(last["octave"])):
        if not last["valid"][i]:
            continue
        o = int(last["octave"][i])
        radius = np.float32(np.float32(th) * F.scale[o])
        if forward:
            cand = F.features_in_area(last["u"][i], last["v"][i], radius, o)
        elif backward:
            cand = F.features_in_area(last["u"][i], last["v"][i], radius, 0, o)
        else:
            cand = F.features_in_area(last["u"][i], last["v"][i], radius, o - 1, o + 1)
        if not cand:
            continue
        best, bi = 256, -1
        for i2 in cand:
            if occupied[i2]:
                continue
            if F.uright[i2] > 0:
                er = abs(np.float32(last["ur"][i]) - np.float32(F.uright[i2]))
                if er > radius:
                    continue
            d = hamming(last["desc"][i], F.desc[i2])
            if d < best:
                best, bi = d, i2
        if best <= th_high:
            assigned[bi] = i
            occupied[bi] = bool(last["has_obs"][i])
            nm += 1
            if check_ori:
                rot[_rot_bin(last["angle"][i], F.kps[bi, 3])].append(bi)
    if check_ori:
        def clear(j):
            assigned[j] = -1
            if cleared is not None:
                cleared.add(j)
        nm -= _rot_cleanup(rot, clear)
    return nm, assigned


def search_by_projection_cur_kf(F, q, th, orb_dist, check_ori=True, cleared=None):
    """src/ORBmatcher.cc:1474-1601. F.occupied[kp]: mvpMapPoints[kp] set. q: valid, u, v, level, angle, desc."""
    assigned = np.full(len(F.kps), -1, np.int32)
    occupied = F.occupied.copy()
    rot = [[] for _ in range(30)]
    nm = 0
    for i in range(len(q["level"])):
        if not q["valid"][i]:
            continue
        lvl = int(q["level"][i])
        radius = np.float32(np.float32(th) * F.scale[lvl])
        cand = F.features_in_area(q["u"][i], q["v"][i], radius, lvl - 1, lvl + 1)
        if not cand:
            continue
        best, bi = 256, -1
        for i2 in cand:
            if occupied[i2]:
                continue
            d = hamming(q["desc"][i], F.desc[i2])
            if d < best:
                best, bi = d, i2
        if best <= orb_dist:
            assigned[bi] = i
            occupied[bi] = True
            nm += 1
            if check_ori:
                rot[_rot_bin(q["angle"][i], F.kps[bi, 3])].append(bi)
    if check_ori:
        def clear(j):
            assigned[j] = -1
            if cleared is not None:
                cleared.add(j)
        nm -= _rot_cleanup(rot, clear)
    return nm, assigned


def search_by_projection_kf_sim3(F, matched, q, th, th_low=50):
    """src/ORBmatcher.cc:292-405."""
    matched = np.asarray(matched, bool).copy()
    assigned = np.full(len(F.kps), -1, np.int32)
    nm = 0
    for i in range(len(q["level"])):
        if not q["valid"][i]:
            continue
        lvl = int(q["level"][i])
        radius = np.float32(np.float32(int(th)) * F.scale[lvl])
        cand = F.features_in_area(q["u"][i], q["v"][i], radius)
        if not cand:
            continue
        best, bi = 256, -1
        for idx in cand:
            if matched[idx]:
                continue
            kl = int(F.kps[idx, 5])
            if kl < lvl - 1 or kl > lvl:
                continue
            d = hamming(q["desc"][i], F.desc[idx])
            if d < best:
                best, bi = d, idx
        if best <= th_low:
            matched[bi] = True
            assigned[bi] = i
            nm += 1
    return nm, assigned


def _walk(fv1, fv2):
    i = j = 0
    while i < len(fv1) and j < len(fv2):
        if fv1[i][0] == fv2[j][0]:
            yield fv1[i][1], fv2[j][1]
            i += 1; j += 1
        elif fv1[i][0] < fv2[j][0]:
            while i < len(fv1) and fv1[i][0] < fv2[j][0]:
                i += 1
        else:
            while j < len(fv2) and fv2[j][0] < fv1[i][0]:
                j += 1


def search_by_bow_kf_f(desc_kf, fv_kf, valid_kf, angle_kf, desc_f, fv_f, angle_f, nnratio=0.7, check_ori=True, th_low=50):
    """src/ORBmatcher.cc:161-290."""
    assigned = np.full(len(desc_f), -1, np.int32)
    rot = [[] for _ in range(30)]
    ratio = np.float32(nnratio)
    nm = 0
    for lkf, lf in _walk(fv_kf, fv_f):
        for ikf in lkf:
            if not valid_kf[ikf]:
                continue
            b1 = b2 = 256
            bi = -1
            for jf in lf:
                if assigned[jf] >= 0:
                    continue
                d = hamming(desc_kf[ikf], desc_f[jf])
                if d < b1:
                    b2, b1, bi = b1, d, jf
                elif d < b2:
                    b2 = d
            if b1 <= th_low:
                if np.float32(b1) < ratio * np.float32(b2):
                    assigned[bi] = ikf
                    if check_ori:
                        rot[_rot_bin(angle_kf[ikf], angle_f[bi])].append(bi)
                    nm += 1
    if check_ori:
        def clear(j):
            assigned[j] = -1
        nm -= _rot_cleanup(rot, clear)
    return nm, assigned


def check_dist_epipolar_line(x1, y1, x2, y2, o2, F12, level_sigma2):
    """src/ORBmatcher.cc:142-159."""
    f = np.float32
    a = f(f(f(x1) * F12[0, 0] + f(y1) * F12[1, 0]) + F12[2, 0])
    b = f(f(f(x1) * F12[0, 1] + f(y1) * F12[1, 1]) + F12[2, 1])
    c = f(f(f(x1) * F12[0, 2] + f(y1) * F12[1, 2]) + F12[2, 2])
    num = f(f(a * f(x2) + b * f(y2)) + c)
    den = f(a * a + b * b)
    if den == 0:
        return False
    dsqr = f(f(num * num) / den)
    return float(dsqr) < 3.84 * float(level_sigma2[o2])


def search_for_triangulation(kf1, kf2, F12, epipole, scale2, sigma2_2, only_stereo=False, check_ori=True, th_low=50):
    """src/ORBmatcher.cc:659-825."""
    f = np.float32
    F12 = np.asarray(F12, f)
    ex, ey = f(epipole[0]), f(epipole[1])
    n1 = len(kf1["desc"])
    m12 = np.full(n1, -1, np.int32)
    matched2 = np.zeros(len(kf2["desc"]), bool)  # never set by the reference either
    rot = [[] for _ in range(30)]
    nm = 0
    for l1, l2 in _walk(kf1["featvec"], kf2["featvec"]):
        for idx1 in l1:
            if kf1["has_mp"][idx1]:
                continue
            st1 = kf1["uright"][idx1] >= 0
            if only_stereo and not st1:
                continue
            best, bi = th_low, -1
            for idx2 in l2:
                if matched2[idx2] or kf2["has_mp"][idx2]:
                    continue
                st2 = kf2["uright"][idx2] >= 0
                if only_stereo and not st2:
                    continue
                d = hamming(kf1["desc"][idx1], kf2["desc"][idx2])
                if d > th_low or d > best:
                    continue
                if not st1 and not st2:
                    dx = f(ex - f(kf2["x"][idx2])); dy = f(ey - f(kf2["y"][idx2]))
                    if f(f(dx * dx) + f(dy * dy)) < f(f(100) * f(scale2[int(kf2["octave"][idx2])])):
                        continue
                if check_dist_epipolar_line(kf1["x"][idx1], kf1["y"][idx1], kf2["x"][idx2], kf2["y"][idx2],
                                            int(kf2["octave"][idx2]), F12, sigma2_2):
                    bi, best = idx2, d
            if bi >= 0:
                m12[idx1] = bi
                nm += 1
                if check_ori:
                    rot[_rot_bin(kf1["angle"][idx1], kf2["angle"][bi])].append(idx1)
    if check_ori:
        def clear(j):
            m12[j] = -1
        nm -= _rot_cleanup(rot, clear)
    return [(i, int(m12[i])) for i in range(n1) if m12[i] >= 0]


def fuse_kf_mappoints(F, inv_sigma2, q, th, th_low=50):
    """The search of src/ORBmatcher.cc:827-977 (bookkeeping 953-972 left to the caller)."""
    f = np.float32
    n = len(q["level"])
    best_idx = np.full(n, -1, np.int32); best_dist = np.full(n, 256, np.int32)
    for i in range(n):
        if not q["valid"][i]:
            continue
        lvl = int(q["level"][i])
        radius = f(f(th) * F.scale[lvl])
        u, v, ur = f(q["u"][i]), f(q["v"][i]), f(q["ur"][i])
        best, bi = 256, -1
        for idx in F.features_in_area(u, v, radius):
            kl = int(F.kps[idx, 5])
            if kl < lvl - 1 or kl > lvl:
                continue
            exx = f(u - F.kps[idx, 0]); eyy = f(v - F.kps[idx, 1])
            if F.uright[idx] >= 0:
                er = f(ur - f(F.uright[idx]))
                e2 = f(f(f(exx * exx) + f(eyy * eyy)) + f(er * er))
                if float(f(e2 * f(inv_sigma2[kl]))) > 7.8:
                    continue
            else:
                e2 = f(f(exx * exx) + f(eyy * eyy))
                if float(f(e2 * f(inv_sigma2[kl]))) > 5.99:
                    continue
            d = hamming(q["desc"][i], F.desc[idx])
            if d < best:
                best, bi = d, idx
        best_dist[i] = best
        if best <= th_low:
            best_idx[i] = bi
    return int((best_idx >= 0).sum()), best_idx, best_dist


def best_in_window(F, q, th, limit):
    """The inner search shared by Fuse(KF,Scw,...) (src/ORBmatcher.cc:1050-1082) and SearchBySim3 (1189-1221, 1269-1301)."""
    n = len(q["level"])
    out = np.full(n, -1, np.int32)
    for i in range(n):
        if not q["valid"][i]:
            continue
        lvl = int(q["level"][i])
        radius = np.float32(np.float32(th) * F.scale[lvl])
        best, bi = 2**31 - 1, -1
        for idx in F.features_in_area(q["u"][i], q["v"][i], radius):
            kl = int(F.kps[idx, 5])
            if kl < lvl - 1 or kl > lvl:
                continue
            d = hamming(q["desc"][i], F.desc[idx])
            if d < best:
                best, bi = d, idx
        if best <= limit:
            out[i] = bi
    return out


def search_by_sim3(F1, F2, q1, q2, th, th_high=100):
    """src/ORBmatcher.cc:1104-1328."""
    m1 = best_in_window(F2, q1, th, th_high)
    m2 = best_in_window(F1, q2, th, th_high)
    m12 = np.full(len(m1), -1, np.int32)
    for i1 in range(len(m1)):
        if m1[i1] >= 0 and m2[m1[i1]] == i1:
            m12[i1] = m1[i1]
    return int((m12 >= 0).sum()), m12


# ---- map point projection (Frame::isInFrustum and the Fuse / Sim3 prologue), scalar restatement -----------
def _mat3_vec(R, x, t):
    """cv::Mat Rcw*P+tcw for 3x3 * 3x1 floats: OpenCV's small-matrix gemm accumulates left to right in float
    (pinned against cv2.gemm in tests/test_projection_oracle.py), then the float matrix add."""
    f = np.float32
    return [f(f(f(f(R[i][0] * x[0]) + f(R[i][1] * x[1])) + f(R[i][2] * x[2])) + t[i]) for i in range(3)]


def _norm3(v):
    """cv::norm of a 3x1 float Mat: squares accumulated in double in order, double sqrt; the caller assigns to float."""
    d = [float(x) for x in v]
    return np.float32(np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]))


def _dot3(a, b):
    """cv::Mat::dot of 3x1 float Mats: products and sum in double, in order."""
    return (float(a[0]) * float(b[0]) + float(a[1]) * float(b[1])) + float(a[2]) * float(b[2])


def predict_scale(max_distance, dist, log_scale_factor, n_levels):
    """MapPoint::PredictScale, src/MapPoint.cc:389-421: ceil(log(ratio)/mfLogScaleFactor), all float (std::log(float))."""
    f = np.float32
    ratio = f(f(max_distance) / f(dist))
    n = int(np.ceil(f(np.log(ratio, dtype=np.float32) / f(log_scale_factor))))
    return 0 if n < 0 else (n_levels - 1 if n >= n_levels else n)


def project_points(cam, pos, normal, max_d, min_d, mode, cos_limit=0.5, th=1.0):
    """mode 0: Frame::isInFrustum, src/Frame.cc:269-325 (+ the radius of src/ORBmatcher.cc:60-67, 133-139);
    mode 1: the prologue of ORBmatcher::Fuse, src/ORBmatcher.cc:849-889 (+ radius th*scale, 892).
    cam: dict Rcw (3x3), tcw, Ow, fx, fy, cx, cy, bf, min_x, max_x, min_y, max_y, log_scale_factor, scale (array)."""
    f = np.float32
    n = len(pos)
    out = dict(alive=np.zeros(n, np.uint8), u=np.zeros(n, f), v=np.zeros(n, f), ur=np.zeros(n, f), level=np.zeros(n, np.int32),
               view_cos=np.zeros(n, f), radius=np.zeros(n, f))
    R = np.asarray(cam["Rcw"], f).reshape(3, 3); t = np.asarray(cam["tcw"], f); Ow = np.asarray(cam["Ow"], f)
    fx, fy, cx, cy, bf = (f(cam[k]) for k in ("fx", "fy", "cx", "cy", "bf"))
    scale = np.asarray(cam["scale"], f)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for i in range(n):
            P = np.asarray(pos[i], f)
            x, y, z = _mat3_vec(R, P, t)
            ok = not (z < f(0))
            invz = f(f(1.0) / z)
            if mode == 0:
                u = f(f(f(fx * x) * invz) + cx); v = f(f(f(fy * y) * invz) + cy)
                ok = ok and not (u < cam["min_x"] or u > cam["max_x"]) and not (v < cam["min_y"] or v > cam["max_y"])
            else:
                u = f(f(fx * f(x * invz)) + cx); v = f(f(fy * f(y * invz)) + cy)
                ok = ok and bool(u >= cam["min_x"] and u < cam["max_x"] and v >= cam["min_y"] and v < cam["max_y"])
            ur = f(u - f(bf * invz))
            maxd, mind = f(f(1.2) * f(max_d[i])), f(f(0.8) * f(min_d[i]))
            PO = [f(P[k] - Ow[k]) for k in range(3)]
            dist = _norm3(PO)
            ok = ok and not (dist < mind or dist > maxd)
            dot = _dot3(PO, np.asarray(normal[i], f))
            view_cos = f(dot / float(dist))
            if mode == 0:
                ok = ok and not (view_cos < f(cos_limit))
            else:
                ok = ok and not (dot < 0.5 * float(dist))
            level, radius = 0, f(0)
            if ok:
                level = predict_scale(max_d[i], dist, cam["log_scale_factor"], len(scale))
                if mode == 0:
                    r = f(2.5) if view_cos > 0.998 else f(4.0)
                    if th != 1.0:
                        r = f(r * f(th))
                    radius = f(r * scale[level])
                else:
                    radius = f(f(th) * scale[level])
            out["alive"][i] = ok; out["u"][i] = u; out["v"][i] = v; out["ur"][i] = ur
            out["level"][i] = level; out["view_cos"][i] = view_cos; out["radius"][i] = radius
    return out


# ---- the projection prologues of the guided searches with their individual quirks (C++ facade parity) ----------
def _project_one(R, P, t, fx, fy, cx, cy, invz_double, uform):
    """Pc = R*P+t; invz as `1.0/z` (double division rounded to float) or `1/z`, `1.0f/z` (float division);
    uform 0: fx*x*invz+cx (src/ORBmatcher.cc:1371-1372), 1: fx*(x*invz)+cx (e.g. 337-341)."""
    f = np.float32
    pc = _mat3_vec(R, P, t)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        invz = f(1.0 / float(pc[2])) if invz_double else f(f(1.0) / pc[2])
        if uform == 0:
            u = f(f(f(fx * pc[0]) * invz) + cx); v = f(f(f(fy * pc[1]) * invz) + cy)
        else:
            u = f(f(fx * f(pc[0] * invz)) + cx); v = f(f(fy * f(pc[1] * invz)) + cy)
    return pc, invz, u, v


def minus_Rt_t(R, t):
    """cv::Mat -R.t()*t (general gemm: double accumulation), e.g. mOw, src/ORBmatcher.cc:1343, 1480."""
    R = np.asarray(R, np.float32).reshape(3, 3).astype(np.float64); t = np.asarray(t, np.float32).astype(np.float64)
    return np.array([-1.0 * ((R[0, i] * t[0] + R[1, i] * t[1]) + R[2, i] * t[2]) for i in range(3)]).astype(np.float32)


def decompose_sim3(Scw):
    """src/ORBmatcher.cc:300-305: scw = sqrt(row0.row0), Rcw = sRcw/scw, tcw = t/scw (cv::Mat / scalar = times the float of the
    double reciprocal), Ow = -Rcw.t()*tcw."""
    f = np.float32
    S = np.asarray(Scw, f).reshape(4, 4)
    scw = f(np.sqrt(_dot3(S[0, :3], S[0, :3])))
    inv = f(1.0 / float(scw))
    R = (S[:3, :3] * inv).astype(f); t = (S[:3, 3] * inv).astype(f)
    return R, t, minus_Rt_t(R, t)


def _in_frame(u, v, b):   # Frame bounds test: u<mnMinX || u>mnMaxX -> out
    return not (u < b[0] or u > b[1]) and not (v < b[2] or v > b[3])


def _in_image(u, v, b):   # KeyFrame::IsInImage, src/KeyFrame.cc:630-633
    return bool(u >= b[0] and u < b[1] and v >= b[2] and v < b[3])


def prologue_cur_last(cam, Tcw_cur, pos, alive):
    """src/ORBmatcher.cc:1361-1380. Returns valid, u, v, ur per point."""
    f = np.float32
    T = np.asarray(Tcw_cur, f).reshape(4, 4)
    n = len(pos)
    out = dict(valid=np.zeros(n, bool), u=np.zeros(n, f), v=np.zeros(n, f), ur=np.zeros(n, f))
    b = (cam["min_x"], cam["max_x"], cam["min_y"], cam["max_y"])
    for i in range(n):
        if not alive[i]:
            continue
        pc, invz, u, v = _project_one(T[:3, :3], np.asarray(pos[i], f), T[:3, 3], cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, 0)
        if invz < 0 or not _in_frame(u, v, b):
            continue
        out["valid"][i] = True; out["u"][i] = u; out["v"][i] = v; out["ur"][i] = f(u - f(f(cam["bf"]) * invz))
    return out


def motion_direction(Tcw_cur, Tcw_last, mb, mono):
    """bForward / bBackward, src/ORBmatcher.cc:1340-1351."""
    f = np.float32
    Tc = np.asarray(Tcw_cur, f).reshape(4, 4); Tl = np.asarray(Tcw_last, f).reshape(4, 4)
    twc = minus_Rt_t(Tc[:3, :3], Tc[:3, 3])
    tlc = _mat3_vec(Tl[:3, :3], twc, Tl[:3, 3])
    return bool(tlc[2] > f(mb) and not mono), bool(-tlc[2] > f(mb) and not mono)


def prologue_scaled(cam, R, t, Ow, pos, normal, max_d, min_d, alive, invz_double, zcheck, uform, bounds, normal_gate, pre=None,
                    dist_from_cam=False):
    """The common shape of src/ORBmatcher.cc:1497-1527 (Cur,KF), 323-363 (KF,Scw), 1007-1047 (Fuse Scw), 1169-1186 (Sim3):
    project, bounds, distance-invariance range, optional viewing-angle gate, PredictScale. pre=(R0,t0): a first camera
    transform (SearchBySim3). Returns valid, u, v, level."""
    f = np.float32
    n = len(pos)
    out = dict(valid=np.zeros(n, bool), u=np.zeros(n, f), v=np.zeros(n, f), level=np.zeros(n, np.int32))
    b = (cam["min_x"], cam["max_x"], cam["min_y"], cam["max_y"])
    for i in range(n):
        if not alive[i]:
            continue
        P = np.asarray(pos[i], f)
        Pin = np.array(_mat3_vec(pre[0], P, pre[1]), f) if pre is not None else P
        pc, invz, u, v = _project_one(R, Pin, t, cam["fx"], cam["fy"], cam["cx"], cam["cy"], invz_double, uform)
        if zcheck and pc[2] < 0:
            continue
        if not (_in_frame(u, v, b) if bounds == "frame" else _in_image(u, v, b)):
            continue
        if dist_from_cam:
            PO = pc
        else:
            PO = [f(P[k] - Ow[k]) for k in range(3)]
        dist = _norm3(PO)
        if dist < f(f(0.8) * f(min_d[i])) or dist > f(f(1.2) * f(max_d[i])):
            continue
        if normal_gate and _dot3(PO, np.asarray(normal[i], f)) < 0.5 * float(dist):
            continue
        out["valid"][i] = True; out["u"][i] = u; out["v"][i] = v
        out["level"][i] = predict_scale(max_d[i], dist, cam["log_scale_factor"], len(cam["scale"]))
    return out


# ---- keyframe database scoring (src/KeyFrameDatabase.cc + DBoW2 L1Scoring), scalar restatement ------------------
def l1_score(id1, w1, id2, w2):
    """L1Scoring::score, Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-66: merge walk over ascending word ids."""
    i = j = 0
    s = 0.0
    while i < len(id1) and j < len(id2):
        if id1[i] == id2[j]:
            vi, wi = float(w1[i]), float(w2[j])
            s += abs(vi - wi) - abs(vi) - abs(wi)
            i += 1; j += 1
        elif id1[i] < id2[j]:
            i += 1
        else:
            j += 1
    return -s / 2.0


def ref_l1_score(id1, w1, id2, w2):
    """The vendored DBoW2 itself (oracle/_ref/libref_dbow.so)."""
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_dbow.so"))
    L.refv_l1_score.restype = C.c_double
    L.refv_l1_score.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    a, b = np.ascontiguousarray(id1, np.int32), np.ascontiguousarray(id2, np.int32)
    x, y = np.ascontiguousarray(w1, np.float64), np.ascontiguousarray(w2, np.float64)
    return L.refv_l1_score(a.ctypes.data, x.ctypes.data, len(a), b.ctypes.data, y.ctypes.data, len(b))


def detect_loop_candidates(db, alive, q_ids, q_w, min_score, connected, neighbours):
    """KeyFrameDatabase::DetectLoopCandidates, src/KeyFrameDatabase.cc:76-197, on a list of BowVectors db[slot] = (ids, weights)
    added in slot order; connected: set of slots; neighbours(slot): GetBestCovisibilityKeyFrames(10) of that keyframe.
    Returns the candidate slots in the reference's output order."""
    inverted = {}
    for slot, (ids, _) in enumerate(db):
        if alive[slot]:
            for w in ids:
                inverted.setdefault(int(w), []).append(slot)
    sharing, words, queried = [], {}, set()
    for w in q_ids:
        for slot in inverted.get(int(w), []):
            if slot not in queried:
                words[slot] = 0
                if slot not in connected:
                    queried.add(slot)
                    sharing.append(slot)
            words[slot] += 1
    if not sharing:
        return []
    max_common = max(words[s] for s in sharing)
    min_common = int(np.float32(max_common) * np.float32(0.8))
    score, matches = {}, []
    for slot in sharing:
        if words[slot] > min_common:
            si = np.float32(l1_score(q_ids, q_w, db[slot][0], db[slot][1]))
            score[slot] = si
            if si >= np.float32(min_score):
                matches.append((si, slot))
    if not matches:
        return []
    acc, best_acc = [], np.float32(min_score)
    for si, slot in matches:
        best_score, acc_score, best_kf = si, si, slot
        for nb in neighbours(slot):
            if nb in queried and words[nb] > min_common:
                acc_score = np.float32(acc_score + score[nb])
                if score[nb] > best_score:
                    best_kf, best_score = nb, score[nb]
        acc.append((acc_score, best_kf))
        if acc_score > best_acc:
            best_acc = acc_score
    retain = np.float32(np.float32(0.75) * best_acc)
    out, seen = [], set()
    for a, kf in acc:
        if a > retain and kf not in seen:
            out.append(kf); seen.add(kf)
    return out


def _sharing(db, alive, q_ids, skip=()):
    """Keyframes sharing a word with the query in the order the reference meets them (ascending word, then insertion order of
    the word's list), with their common-word counts. skip: slots left out entirely."""
    inverted = {}
    for slot, (ids, _) in enumerate(db):
        if alive[slot]:
            for w in ids:
                inverted.setdefault(int(w), []).append(slot)
    order, words = [], {}
    for w in q_ids:
        for slot in inverted.get(int(w), []):
            if slot in skip:
                continue
            if slot not in words:
                words[slot] = 0
                order.append(slot)
            words[slot] += 1
    return order, words


def detect_covisibility_candidates(db, alive, q_ids, q_w, min_score, ignore, neighbours, covis_score):
    """KeyFrameDatabase::DetectCovisibilityCandidates, src/KeyFrameDatabase.cc:199-308 (the fork's place recognition for
    MapFusion::CovisibilityDiscovery). Unlike DetectLoopCandidates it never stores the similarity in the keyframe
    (mCovisScore is only READ, 270-276), so the covisibility accumulation adds whatever the keyframes held before:
    covis_score[slot] is that state, an explicit input here."""
    f = np.float32
    order, words = _sharing(db, alive, q_ids, set(ignore))
    if not order:
        return []
    min_common = int(f(max(words[s] for s in order)) * f(0.8))
    matches = []
    for slot in order:
        if words[slot] > min_common:
            si = f(l1_score(q_ids, q_w, db[slot][0], db[slot][1]))
            if si >= f(min_score):
                matches.append((si, slot))
    if not matches:
        return []
    acc, best_acc = [], f(min_score)
    for si, slot in matches:
        best_score, acc_score, best_kf = si, si, slot
        for nb in neighbours(slot):
            if nb in words and words[nb] > min_common:
                acc_score = f(acc_score + f(covis_score[nb]))
                if f(covis_score[nb]) > best_score:
                    best_kf, best_score = nb, f(covis_score[nb])
        acc.append((acc_score, best_kf))
        if acc_score > best_acc:
            best_acc = acc_score
    retain = f(f(0.75) * best_acc)
    out, seen = [], set()
    for a, kf in acc:
        if a > retain and kf not in seen:
            out.append(kf); seen.add(kf)
    return out


def detect_relocalization_candidates(db, alive, q_ids, q_w, neighbours, reloc_score):
    """KeyFrameDatabase::DetectRelocalizationCandidates, src/KeyFrameDatabase.cc:310-420. reloc_score[slot] = the keyframes'
    mRelocScore, updated in place: a neighbour that shares a word but stays below the common-word floor contributes the score
    of an EARLIER query (378-387 test only mnRelocQuery)."""
    f = np.float32
    order, words = _sharing(db, alive, q_ids)
    if not order:
        return []
    min_common = int(f(max(words[s] for s in order)) * f(0.8))
    matches = []
    for slot in order:
        if words[slot] > min_common:
            si = f(l1_score(q_ids, q_w, db[slot][0], db[slot][1]))
            reloc_score[slot] = si
            matches.append((si, slot))
    if not matches:
        return []
    acc, best_acc = [], f(0)
    for si, slot in matches:
        best_score, acc_score, best_kf = si, si, slot
        for nb in neighbours(slot):
            if nb not in words:
                continue
            acc_score = f(acc_score + f(reloc_score[nb]))
            if f(reloc_score[nb]) > best_score:
                best_kf, best_score = nb, f(reloc_score[nb])
        acc.append((acc_score, best_kf))
        if acc_score > best_acc:
            best_acc = acc_score
    retain = f(f(0.75) * best_acc)
    out, seen = [], set()
    for a, kf in acc:
        if a > retain and kf not in seen:
            out.append(kf); seen.add(kf)
    return out
