"""One tensor-core kNN-2 call (100k x 100k) for ncu."""
import ctypes as C

import numpy as np
import torch

from multiagent_orb_slam2_b200 import _lib

L = _lib.lib()
vp = C.c_void_p
L.orbm_knn2_mma_device.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp]
dev = torch.device("cuda", 0)
n = 100000
rng = np.random.default_rng(0)
A = torch.as_tensor(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
B = torch.as_tensor(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
out = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)]
for _ in range(2):
    rc = L.orbm_knn2_mma_device(vp(A.data_ptr()), n, vp(B.data_ptr()), n, vp(out[0].data_ptr()), vp(out[1].data_ptr()), vp(out[2].data_ptr()), None)
    assert rc == 0
torch.cuda.synchronize()
print("ok", int(out[1].min()), int(out[1].max()))
