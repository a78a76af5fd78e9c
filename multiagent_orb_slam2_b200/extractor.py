"""Host-side mirror of ORB_SLAM2::ORBextractor (/root/reference/include/ORBextractor.h:45-111) over
the C ABI. Same constructor arguments, same call semantics (image, mask ignored -> keypoints,
descriptors), same getters; `mvImagePyramid` is a lazily downloaded view of the device pyramid."""
import ctypes as C

import numpy as np

from . import _lib

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])


class ORBextractor:
    HARRIS_SCORE, FAST_SCORE = 0, 1  # ORBextractor.h:49 (unused by the reference too)

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width=None, height=None, device=0,
                 max_batch=1):
        self._L = _lib.lib()
        self.cfg = _lib.Config(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST))
        self.device, self.max_batch = int(device), int(max_batch)
        self._h = None
        self._shape = None
        if width is not None:
            self._create(int(width), int(height))

    # -- lifetime ---------------------------------------------------------------------------------
    def _create(self, w, h):
        self.close()
        hd = C.c_void_p()
        _lib.check(self._L.orbx_create(C.byref(self.cfg), self.device, w, h, self.max_batch, C.byref(hd)))
        self._h, self._shape = hd, (h, w)
        self.cap = self._L.orbx_max_keypoints(hd)

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- getters (ORBextractor.h:63-83) -----------------------------------------------------------
    def _tables(self):
        n = self.cfg.nlevels
        if self._h is None:
            raise _lib.OrbError(_lib.ORB_EINVAL, "extractor has no geometry yet (call it on an image first)")
        t = [np.empty(n, np.float32) for _ in range(4)] + [np.empty(n, np.int32)]
        _lib.check(self._L.orbx_get_tables(self._h, *[a.ctypes.data_as(C.c_void_p) for a in t]))
        return t

    def GetLevels(self):
        return self.cfg.nlevels

    def GetScaleFactor(self):
        return float(np.float32(self.cfg.scale_factor))

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        return self._tables()[4]

    # -- operator() ---------------------------------------------------------------------------------
    def __call__(self, image, mask=None):
        """Returns (keypoints, descriptors): a structured array of KP_DTYPE and an (n,32) uint8
        array. An empty image returns empty outputs (the reference's silent no-op)."""
        if image is None or image.size == 0:
            return np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        kps, desc, counts = self.extract_batch(image[None])
        n = int(counts[0])
        return kps[0, :n].copy(), desc[0, :n].copy()

    def extract_batch(self, images):
        """images: (n, H, W) uint8, C-contiguous rows. Returns padded (n, cap) keypoints,
        (n, cap, 32) descriptors and the per-frame counts."""
        assert images.dtype == np.uint8 and images.ndim == 3 and images.strides[2] == 1
        n, h, w = images.shape
        if self._h is None or self._shape != (h, w):
            self._create(w, h)
        if n > self.max_batch:
            raise _lib.OrbError(_lib.ORB_EINVAL, "batch of %d frames exceeds max_batch=%d" % (n, self.max_batch))
        if images.strides[0] != images.strides[1] * h and n > 1:
            images = np.ascontiguousarray(images)
        kps = np.empty((n, self.cap), KP_DTYPE)
        desc = np.empty((n, self.cap, 32), np.uint8)
        counts = np.empty(n, np.int32)
        _lib.check(self._L.orbx_extract_batch(self._h, C.c_void_p(images.ctypes.data), images.strides[1], images.strides[0], n,
                                              kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), self.cap,
                                              counts.ctypes.data_as(C.c_void_p)))
        return kps, desc, counts

    def extract_device(self, d_ptr, pitch, frame_stride, n, stream=0):
        """Frames already resident on the device; enqueues only. Results stay on the device."""
        _lib.check(self._L.orbx_extract_device(self._h, C.c_void_p(d_ptr), pitch, frame_stride, n, C.c_void_p(stream)))

    def device_results(self):
        k, d, c, cap = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
        _lib.check(self._L.orbx_device_results(self._h, C.byref(k), C.byref(d), C.byref(c), C.byref(cap)))
        return k.value, d.value, c.value, cap.value

    # -- mvImagePyramid (ORBextractor.h:85) and stage taps -----------------------------------------
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _lib.check(self._L.orbx_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _lib.check(self._L.orbx_pyramid_level(self._h, frame, level, out.ctypes.data_as(C.c_void_p), w))
        return out

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.cfg.nlevels)]

    def blurred_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _lib.check(self._L.orbx_debug_blurred_level(self._h, frame, level, out.ctypes.data_as(C.c_void_p), w))
        return out

    def candidates(self, level, frame=0):
        cap = 1 << 17
        out = np.empty((cap, 3), np.int32)
        n = C.c_int()
        _lib.check(self._L.orbx_debug_candidates(self._h, frame, level, out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        return out[:n.value].copy()


def quadtree(xs, ys, scores, minX, maxX, minY, maxY, N, device=0):
    """Stand-alone run of the distribution kernel on a host candidate list (parity tests)."""
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); sc = np.ascontiguousarray(scores, np.int32)
    cap = max(N, 64) + 64
    out = np.empty(cap, np.int32)
    n = C.c_int()
    L = _lib.lib()
    _lib.check(L.orbx_debug_quadtree(device, xs.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                     len(xs), minX, maxX, minY, maxY, N, out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    return out[:n.value].copy()


def compute_stereo_matches(left, right, mbf, mb, frame=0):
    """Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:466-640) on the last results of two
    ORBextractor objects (left, right image of a rectified pair). Returns (mvuRight, mvDepth, kept):
    float32 arrays over the left keypoints (-1 = no match)."""
    L = _lib.lib()
    cap = left.cap
    u = np.empty(cap, np.float32); d = np.empty(cap, np.float32)
    kept = C.c_int()
    _lib.check(L.orbm_stereo_match(left._h, right._h, frame, float(mbf), float(mb), u.ctypes.data_as(C.c_void_p),
                                   d.ctypes.data_as(C.c_void_p), cap, C.byref(kept)))
    return u, d, kept.value


def compute_stereo_from_rgbd(ex, depth_image, mbf, frame=0, d_kps_un=None):
    """Frame::ComputeStereoFromRGBD (/root/reference/src/Frame.cc:643-664) on the device-resident keypoints of the last call
    of `ex`: depth_image float32 H x W (metres, already scaled by mDepthMapFactor as in Tracking::GrabImageRGBD).
    d_kps_un: device pointer to mvKeysUn of the frame (DeviceFrameGrid(..., K=, dist_coef=).d_kps_un) for a distorted camera -
    the depth is sampled at the raw position, the right coordinate uses the undistorted x. Returns (mvuRight, mvDepth) over
    the frame's keypoints."""
    import torch
    L = _lib.lib()
    kp, _, cn, cap = ex.device_results()
    dev = torch.device("cuda", ex.device)
    img = torch.as_tensor(np.ascontiguousarray(depth_image, np.float32)).to(dev)
    u = torch.empty(cap, dtype=torch.float32, device=dev); d = torch.empty(cap, dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(L.orbm_stereo_from_rgbd_device(ex.device, C.c_void_p(kp + frame * cap * 24), C.c_void_p(d_kps_un) if d_kps_un else None,
                                              C.c_void_p(cn + frame * 4), cap,
                                              C.c_void_p(img.data_ptr()), img.shape[1], img.shape[0], img.shape[1] * 4, float(mbf),
                                              C.c_void_p(u.data_ptr()), C.c_void_p(d.data_ptr()), st))
    torch.cuda.synchronize(dev)
    return u.cpu().numpy(), d.cpu().numpy()
