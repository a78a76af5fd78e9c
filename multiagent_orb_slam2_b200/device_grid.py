"""Device-resident Frame grid + windowed search: the keypoints and descriptors an extractor
left in HBM are bucketed into the reference Frame's 64 x 48 grid on the device (Frame::AssignFeaturesToGrid,
/root/reference/src/Frame.cc:230-245) and queried without a host round trip (Frame::GetFeaturesInArea 327-380 fused with
the best / second Hamming loop). torch is plumbing (device arrays for the queries and results)."""
import ctypes as C

import numpy as np
import torch

from . import _lib


class DeviceFrameGrid:
    def __init__(self, extractor, frame=0, min_x=0.0, min_y=0.0, max_x=None, max_y=None, stream=None, K=None, dist_coef=None):
        """extractor: an ORBextractor whose last call produced `frame`; bounds default to the image
        (mnMinX.. of an undistorted / rectified camera, src/Frame.cc:457-462).

        Distorted cameras (TUM RGB-D, EuRoC mono): pass K = (fx, fy, cx, cy) and dist_coef = (k1, k2, p1, p2[, k3]). The grid
        is then built over mvKeysUn (Frame::UndistortKeyPoints on the device, self.d_kps_un) inside the undistorted image
        bounds (Frame::ComputeImageBounds), exactly like the reference's Frame; without them the raw keypoints are used,
        which is only right for a camera without distortion.

        Stream ordering: the grid is built on `stream` (a raw cudaStream_t / torch.cuda.Stream; default: torch's current
        stream on the extractor's device - never the legacy NULL stream, which the library's non-blocking streams are not
        ordered against). The extraction that produced the keypoints must already be ordered before that stream: the
        synchronous ORBextractor calls are; after an asynchronous launch (frontend.process_async) pass the stream the
        extraction ran on, or make `stream` wait on its event first. Queries issued on any other stream wait for the
        build through an event recorded here."""
        self.L = _lib.lib()
        self.ex = extractor
        h, w = extractor._shape
        kp, de, cn, cap = extractor.device_results()
        self.cap = cap
        self.d_kps = kp + frame * cap * 24
        self.d_desc = de + frame * cap * 32
        self.d_count = cn + frame * 4
        self.dev = torch.device("cuda", extractor.device)
        g = C.c_void_p()
        _lib.check(self.L.orbm_grid_create(extractor.device, cap, C.byref(g)))
        self._g = g
        self.d_kps_un = self.d_kps
        self.bounds = (float(min_x), float(w if max_x is None else max_x), float(min_y), float(h if max_y is None else max_y))
        with torch.cuda.device(self.dev):
            if stream is None:
                ts = torch.cuda.current_stream(self.dev)
            elif isinstance(stream, torch.cuda.Stream):
                ts = stream
            else:
                ts = torch.cuda.ExternalStream(int(stream), device=self.dev)
            if dist_coef is not None and float(np.float32(dist_coef[0])) != 0.0:
                if K is None:
                    raise ValueError("dist_coef needs the camera matrix K = (fx, fy, cx, cy)")
                d = np.ascontiguousarray(dist_coef, np.float32)
                fx, fy, cx, cy = (float(np.float32(v)) for v in K)
                b = np.zeros(4, np.float32)
                _lib.check(self.L.orbm_image_bounds(w, h, fx, fy, cx, cy, d.ctypes.data_as(C.c_void_p), len(d), b.ctypes.data_as(C.c_void_p)))
                self.bounds = tuple(float(v) for v in b)
                with torch.cuda.stream(ts):
                    self._un = torch.empty(cap * 24, dtype=torch.uint8, device=self.dev)
                self.d_kps_un = self._un.data_ptr()
                _lib.check(self.L.orbm_undistort_keypoints_device(extractor.device, C.c_void_p(self.d_kps), C.c_void_p(self.d_count), cap, fx, fy,
                                                                  cx, cy, d.ctypes.data_as(C.c_void_p), len(d), C.c_void_p(self.d_kps_un),
                                                                  C.c_void_p(ts.cuda_stream)))
            _lib.check(self.L.orbm_grid_build_device(g, C.c_void_p(self.d_kps_un), C.c_void_p(self.d_count), self.bounds[0], self.bounds[2],
                                                     self.bounds[1], self.bounds[3], C.c_void_p(ts.cuda_stream)))
            self._built = torch.cuda.Event()
            self._built.record(ts)

    def _query_stream(self):
        """torch's current stream, ordered after the grid build."""
        st = torch.cuda.current_stream(self.dev)
        st.wait_event(self._built)
        return C.c_void_p(st.cuda_stream)

    def undistorted_keypoints(self, n):
        """mvKeysUn of the frame as a structured host array (first n keypoints)."""
        from .extractor import KP_DTYPE
        torch.cuda.synchronize(self.dev)
        view = _raw_device_bytes(self.d_kps_un, n * 24, self.dev)
        return view.cpu().numpy().view(KP_DTYPE)

    def close(self):
        if getattr(self, "_g", None):
            self.L.orbm_grid_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def window_knn2(self, query_desc, x, y, r, min_level=-1, max_level=-1):
        """For every query (descriptor, window centre x/y, radius r, level window): best / second Hamming distance
        among the frame's keypoints inside the window, index of the first minimum in GetFeaturesInArea order, and
        the octaves of best / second. Arrays in, numpy arrays out."""
        n = len(query_desc)
        t = lambda a, dt: torch.as_tensor(np.broadcast_to(np.asarray(a, dt), (n,)).copy()).to(self.dev)
        q = torch.as_tensor(np.ascontiguousarray(query_desc, np.uint8)).to(self.dev)
        tx, ty, tr = t(x, np.float32), t(y, np.float32), t(r, np.float32)
        tmin, tmax = t(min_level, np.int32), t(max_level, np.int32)
        out = [torch.empty(n, dtype=torch.int32, device=self.dev) for _ in range(5)]
        p = lambda z: C.c_void_p(z.data_ptr())
        st = self._query_stream()
        _lib.check(self.L.orbm_window_knn2_device(self._g, C.c_void_p(self.d_desc), p(q), n, p(tx), p(ty), p(tr), p(tmin), p(tmax),
                                                  p(out[0]), p(out[1]), p(out[2]), p(out[3]), p(out[4]), st))
        torch.cuda.synchronize(self.dev)
        return [o.cpu().numpy() for o in out]

    def window_lists(self, query_desc, x, y, r, min_level=-1, max_level=-1, cap=None):
        """CSR candidate lists (offsets, cands, dist) of the windows, candidates in GetFeaturesInArea order with
        their Hamming distances - input of the host-side replay of the stateful searches."""
        n = len(query_desc)
        t = lambda a, dt: torch.as_tensor(np.broadcast_to(np.asarray(a, dt), (n,)).copy()).to(self.dev)
        q = torch.as_tensor(np.ascontiguousarray(query_desc, np.uint8)).to(self.dev)
        tx, ty, tr = t(x, np.float32), t(y, np.float32), t(r, np.float32)
        tmin, tmax = t(min_level, np.int32), t(max_level, np.int32)
        cap = cap or 64 * n
        p = lambda z: C.c_void_p(z.data_ptr())
        st = self._query_stream()
        while True:
            offs = torch.empty(n + 1, dtype=torch.int32, device=self.dev)
            cands = torch.empty(cap, dtype=torch.int32, device=self.dev)
            dist = torch.empty(cap, dtype=torch.int16, device=self.dev)
            total = C.c_int()
            rc = self.L.orbm_window_lists_device(self._g, C.c_void_p(self.d_desc), p(q), n, p(tx), p(ty), p(tr), p(tmin), p(tmax),
                                                 p(offs), p(cands), p(dist), cap, C.byref(total), st)
            if rc == _lib.ORB_ECAPACITY:
                cap = total.value
                continue
            _lib.check(rc)
            break
        torch.cuda.synchronize(self.dev)
        return offs.cpu().numpy(), cands[:total.value].cpu().numpy(), dist[:total.value].cpu().numpy()


def _raw_device_bytes(ptr, nbytes, dev):
    class _Ext:
        pass
    e = _Ext()
    e.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    with torch.cuda.device(dev):
        return torch.as_tensor(e, device=dev)[:nbytes]
