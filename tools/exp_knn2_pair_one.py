"""One size of the CTA-pair / one-CTA tensor-core matcher for ncu: python tools/exp_knn2_pair_one.py <backend> <na> <nb>."""
import ctypes as C
import sys

import numpy as np
import torch

from multiagent_orb_slam2_b200 import _lib

L = _lib.lib()
vp = C.c_void_p
backend, na, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
torch.cuda.init(); torch.zeros(1, device=dev)
a, b = C.c_int(), C.c_int()
_lib.check(L.orbm_debug_mma_occupancy(C.byref(a), C.byref(b)))
print("resident CTA pairs on the GPU:", a.value, " one-CTA blocks per SM:", b.value)
rng = np.random.default_rng(1)
A = torch.as_tensor(rng.integers(0, 256, (na, 32), dtype=np.uint8)).to(dev)
B = torch.as_tensor(rng.integers(0, 256, (nb, 32), dtype=np.uint8)).to(dev)
out = [torch.empty(na, dtype=torch.int32, device=dev) for _ in range(3)]
_lib.check(L.orbm_set_knn2_backend(backend))
for _ in range(2):
    _lib.check(L.orbm_knn2_device(vp(A.data_ptr()), na, vp(B.data_ptr()), nb, *[vp(o.data_ptr()) for o in out], None))
torch.cuda.synchronize()
print("ok", int(out[1].sum()))
