"""One 128 x 256 tile through expand + TMA (SWIZZLE_128B) + tcgen05.mma kind::i8 + tcgen05.ld, against numpy."""
import ctypes as C
import sys

import numpy as np

from multiagent_orb_slam2_b200 import _lib

L = _lib.lib()
rng = np.random.default_rng(0)
A = rng.integers(0, 256, (128, 32), dtype=np.uint8)
B = rng.integers(0, 256, (256, 32), dtype=np.uint8)
B[5] = A[7]; B[200] = ~A[100]
out = np.zeros((128, 256), np.int32)
L.orbm_debug_mma_dot.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
rc = L.orbm_debug_mma_dot(0, A.ctypes.data, B.ctypes.data, out.ctypes.data)
if rc:
    print("rc", rc, _lib.last_error() if hasattr(_lib, "last_error") else L.orb_last_error())
    sys.exit(1)
ham = np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(2).astype(np.int32)
want = 256 - 2 * ham
print("equal:", np.array_equal(out, want), "mismatches:", int((out != want).sum()), "of", out.size)
if not np.array_equal(out, want):
    bad = np.argwhere(out != want)
    print("first bad", bad[:5], out[tuple(bad[0])], want[tuple(bad[0])])
    print("rows with errors", np.unique(bad[:, 0])[:20], "cols", np.unique(bad[:, 1])[:20])
    print(out[:4, :8]); print(want[:4, :8])
