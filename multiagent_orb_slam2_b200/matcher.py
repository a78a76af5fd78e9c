"""Host-side mirror of ORB_SLAM2::ORBmatcher (/root/reference/include/ORBmatcher.h:37-102) over the
C ABI, on plain arrays: the reference's searches walk Frame/KeyFrame/MapPoint pointer graphs (out of
scope); what they compute on descriptors is exposed here with the same
constants, thresholds and acceptance rules."""
import ctypes as C

import numpy as np

from . import _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _desc(a):
    a = np.ascontiguousarray(a, np.uint8)
    assert a.ndim == 2 and a.shape[1] == 32
    return a


class ORBmatcher:
    TH_HIGH = 100      # src/ORBmatcher.cc:37
    TH_LOW = 50        # :38
    HISTO_LENGTH = 30  # :39

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio = float(np.float32(nnratio))
        self.mbCheckOrientation = bool(checkOri)
        self.device = int(device)
        self._L = _lib.lib()

    @staticmethod
    def DescriptorDistance(a, b):
        """Host inline, as in the reference (src/ORBmatcher.cc:1649-1665)."""
        a = np.ascontiguousarray(a, np.uint8).reshape(32); b = np.ascontiguousarray(b, np.uint8).reshape(32)
        return int(np.unpackbits(a ^ b).sum())

    # -- device primitives on host arrays -----------------------------------------------------------
    def knn2(self, A, B):
        """Best / second-best Hamming distance and first-minimum index of every row of A in B."""
        A, B = _desc(A), _desc(B)
        idx = np.empty(len(A), np.int32); d1 = np.empty(len(A), np.int32); d2 = np.empty(len(A), np.int32)
        _lib.check(self._L.orbm_knn2(self.device, _p(A), len(A), _p(B), len(B), _p(idx), _p(d1), _p(d2)))
        return idx, d1, d2

    def knn2_lists(self, A, B, offsets, cands):
        A, B = _desc(A), _desc(B)
        offsets = np.ascontiguousarray(offsets, np.int32); cands = np.ascontiguousarray(cands, np.int32)
        assert len(offsets) == len(A) + 1
        idx = np.empty(len(A), np.int32); d1 = np.empty(len(A), np.int32); d2 = np.empty(len(A), np.int32)
        _lib.check(self._L.orbm_knn2_lists(self.device, _p(A), len(A), _p(B), len(B), _p(offsets), _p(cands), _p(idx), _p(d1), _p(d2)))
        return idx, d1, d2

    def distance_matrix(self, A, B):
        A, B = _desc(A), _desc(B)
        out = np.empty((len(A), len(B)), np.int16)
        _lib.check(self._L.orbm_distance_matrix(self.device, _p(A), len(A), _p(B), len(B), _p(out)))
        return out

    # -- acceptance rules ---------------------------------------------------------------------------
    def accept(self, idx, d1, d2, th=None, inclusive=False):
        """match[i] = idx[i] if d1 (< | <=) th and float(d1) < nnratio*float(d2) else -1
        (src/ORBmatcher.cc:600-603 strict; 229-232 / 461-463 inclusive)."""
        th = self.TH_LOW if th is None else th
        ok = (d1 <= th) if inclusive else (d1 < th)
        ok &= d1.astype(np.float32) < np.float32(self.mfNNratio) * d2.astype(np.float32)
        ok &= idx >= 0
        return np.where(ok, idx, -1).astype(np.int32)

    def SearchBruteForce(self, A, B, th=None):
        """The SearchByBoW(KF,KF) inner loop with the vocabulary gate removed (BASELINE configs 3, 5):
        ratio-tested nearest neighbour of every row of A in B. Returns match indices (-1 = none)."""
        idx, d1, d2 = self.knn2(A, B)
        return self.accept(idx, d1, d2, th)

    @staticmethod
    def ComputeThreeMaxima(hist_sizes):
        """src/ORBmatcher.cc:1603-1644 on the bin populations; returns (ind1, ind2, ind3)."""
        max1 = max2 = max3 = 0
        i1 = i2 = i3 = -1
        for i, s in enumerate(hist_sizes):
            s = int(s)
            if s > max1:
                max3, max2, max1 = max2, max1, s
                i3, i2, i1 = i2, i1, i
            elif s > max2:
                max3, max2 = max2, s
                i3, i2 = i2, i
            elif s > max3:
                max3, i3 = s, i
        lim = np.float32(0.1) * np.float32(max1)  # 0.1f*(float)max1, evaluated in fp32
        if np.float32(max2) < lim:
            i2 = i3 = -1
        elif np.float32(max3) < lim:
            i3 = -1
        return i1, i2, i3
