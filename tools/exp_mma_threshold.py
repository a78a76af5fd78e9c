"""Where does the tensor-core matcher overtake the POPC kernel? Square problems and batches of frame pairs."""
import ctypes as C

import numpy as np
import torch

from multiagent_orb_slam2_b200 import _lib

L = _lib.lib()
vp = C.c_void_p
dev = torch.device("cuda", 0)


def bench(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(0)
for n in (1000, 1500, 2013, 3000, 4096, 6000, 8000):
    A = torch.as_tensor(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
    B = torch.as_tensor(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
    out = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)]
    call = lambda: L.orbm_knn2_device(vp(A.data_ptr()), n, vp(B.data_ptr()), n, vp(out[0].data_ptr()), vp(out[1].data_ptr()), vp(out[2].data_ptr()), None)
    res = []
    for backend in (1, 2):
        L.orbm_set_knn2_backend(backend)
        res.append(bench(call))
    print("single %5d x %5d: popc %.4f ms  tensor %.4f ms" % (n, n, res[0], res[1]))
for pairs in (2, 4, 8, 16, 64):
    n = 1000
    S = torch.as_tensor(rng.integers(0, 256, (pairs + 1, n, 32), dtype=np.uint8)).to(dev)
    cnt = torch.full((pairs + 1,), n, dtype=torch.int32, device=dev)
    pr = torch.as_tensor(np.stack([np.arange(pairs), np.arange(pairs) + 1], 1).astype(np.int32)).to(dev)
    out = [torch.empty((pairs, n), dtype=torch.int32, device=dev) for _ in range(3)]
    call = lambda: L.orbm_knn2_pairs_device(vp(S.data_ptr()), vp(cnt.data_ptr()), n, vp(pr.data_ptr()), pairs, vp(out[0].data_ptr()), vp(out[1].data_ptr()),
                                            vp(out[2].data_ptr()), None)
    res = []
    for backend in (1, 2):
        L.orbm_set_knn2_backend(backend)
        res.append(bench(call))
    print("%3d pairs of 1000 x 1000: popc %.4f ms  tensor %.4f ms" % (pairs, res[0], res[1]))
L.orbm_set_knn2_backend(0)
