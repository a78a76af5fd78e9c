"""One agent's frontend on one GPU: ORB extraction of a batch of frames followed by frame-to-frame
brute-force matching of consecutive frames (descriptor sets stay in HBM between the two). torch is
used for device/pinned memory and streams only; all compute goes through the C ABI.

Mirrors what Tracking does per frame in the reference - build a Frame (Frame::ExtractORB,
/root/reference/src/Frame.cc:247-253) and match it against the previous one with the ratio test of
ORBmatcher (src/ORBmatcher.cc:461-463) - with the candidate gate removed (BASELINE config 3)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .extractor import KP_DTYPE, ORBextractor
from .matcher import ORBmatcher


class AgentFrontend:
    def __init__(self, width, height, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
                 device=0, max_batch=64, nnratio=0.9, th=ORBmatcher.TH_LOW):
        self.device = int(device)
        self.dev = torch.device("cuda", self.device)
        self.L = _lib.lib()
        self.ex = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width, height, device, max_batch)
        self.w, self.h, self.B = width, height, max_batch
        self.nnratio, self.th = float(nnratio), int(th)
        self.cap = self.ex.cap
        kp, de, cn, cap = self.ex.device_results()
        self.d_kps, self.d_desc, self.d_counts = kp, de, cn
        with torch.cuda.device(self.dev):
            z = lambda: torch.empty((max_batch, self.cap), dtype=torch.int32, device=self.dev)
            self.idx, self.d1, self.d2, self.match = z(), z(), z(), z()
        self._pinned = None
        self._pairs = {}
        with torch.cuda.device(self.dev):
            self._ev_up, self._ev_done = torch.cuda.Event(), torch.cuda.Event()

    _compute_streams = {}

    @classmethod
    def _compute_stream(cls, device):
        st = cls._compute_streams.get(device)
        if st is None:
            st = cls._compute_streams[device] = torch.cuda.Stream(torch.device("cuda", device))
        return st

    # ---- device-resident path ---------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def match_consecutive(self, n):
        """frame i (queries) against frame (i+1) % n (database), i = 0..n-1, on the current stream."""
        st = C.c_void_p(self._stream())
        cap, L = self.cap, self.L
        pairs = self._pairs.get(n)
        if pairs is None:
            i = torch.arange(n, dtype=torch.int32)
            pairs = self._pairs[n] = torch.stack([i, (i + 1) % n], 1).contiguous().to(self.dev)
        _lib.check(L.orbm_knn2_pairs_device(C.c_void_p(self.d_desc), C.c_void_p(self.d_counts), cap, C.c_void_p(pairs.data_ptr()), n,
                                            C.c_void_p(self.idx.data_ptr()), C.c_void_p(self.d1.data_ptr()),
                                            C.c_void_p(self.d2.data_ptr()), st))
        _lib.check(L.orbm_ratio_filter_device(C.c_void_p(self.idx.data_ptr()), C.c_void_p(self.d1.data_ptr()), C.c_void_p(self.d2.data_ptr()),
                                              n * cap, self.th, 1, self.nnratio, C.c_void_p(self.match.data_ptr()), st))

    def process_device(self, d_images):
        """d_images: uint8 CUDA tensor (n, H, W), rows contiguous. Enqueues extraction + matching on
        the current stream; results stay in HBM (self.match, handle-owned keypoints/descriptors)."""
        n = d_images.shape[0]
        self.ex.extract_device(d_images.data_ptr(), d_images.stride(1), d_images.stride(0), n, self._stream())
        self.match_consecutive(n)

    # ---- host-buffer (end-to-end) path --------------------------------------------------------------
    def pinned_outputs(self):
        if self._pinned is None:
            B, cap = self.B, self.cap
            p = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
            self._pinned = dict(kps=p((B, cap, 24), torch.uint8), desc=p((B, cap, 32), torch.uint8),
                                counts=p((B,), torch.int32), match=p((B, cap), torch.int32))
        return self._pinned

    def process_async(self, host_images, out=None):
        """host_images: (n, H, W) uint8 numpy array or CPU tensor (pinned for full overlap). Enqueues
        H2D, extraction, matching and the D2H of keypoints, descriptors, counts and matches on the
        current stream; synchronise the stream before reading `out`."""
        out = out or self.pinned_outputs()
        if isinstance(host_images, torch.Tensor):
            ptr, s0, s1, n = host_images.data_ptr(), host_images.stride(0), host_images.stride(1), host_images.shape[0]
        else:
            ptr, s0, s1, n = host_images.ctypes.data, host_images.strides[0], host_images.strides[1], host_images.shape[0]
        cur = torch.cuda.current_stream(self.dev)
        st = C.c_void_p(cur.cuda_stream)
        h = self.ex._h
        _lib.check(self.L.orbx_upload_frames(h, C.c_void_p(ptr), s1, s0, n, st))
        # The kernels of every agent of this GPU go through ONE compute stream: copies of different agents
        # overlap with it, but kernels of two agents are never co-resident on an SM (measured on B200: with
        # the kernels of two handles interleaved a step takes 3.5 ms instead of 3.1 ms - the shared-memory
        # heavy FAST blocks squeeze the L1 of the gather-bound description kernel).
        cs = self._compute_stream(self.device)
        self._ev_up.record(cur)
        cs.wait_event(self._ev_up)
        with torch.cuda.stream(cs):
            _lib.check(self.L.orbx_extract_staged(h, n, C.c_void_p(cs.cuda_stream)))
            self.match_consecutive(n)
        self._ev_done.record(cs)
        cur.wait_event(self._ev_done)
        _lib.check(self.L.orbx_download_results(h, n, C.c_void_p(out["kps"].data_ptr()), C.c_void_p(out["desc"].data_ptr()), self.cap,
                                                C.c_void_p(out["counts"].data_ptr()), st))
        out["match"][:n].copy_(self.match[:n], non_blocking=True)
        return out

    def process(self, host_images):
        out = self.process_async(host_images)
        torch.cuda.current_stream(self.dev).synchronize()
        n = host_images.shape[0]
        kps = out["kps"].numpy()[:n].view(KP_DTYPE).reshape(n, self.cap)
        return kps, out["desc"].numpy()[:n], out["counts"].numpy()[:n], out["match"].numpy()[:n]

    # bytes moved per frame by process_async (bench.py's e2e accounting)
    def h2d_bytes_per_frame(self):
        return self.w * self.h

    def d2h_bytes_per_frame(self):
        return self.cap * (24 + 32 + 4) + 4
