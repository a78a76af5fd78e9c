"""Tensor-core Hamming kNN-2 (orbm_knn2_mma_device) against the POPC kernel (orbm_knn2_device, oracle-verified): results and speed."""
import ctypes as C
import sys

import numpy as np
import torch

from multiagent_orb_slam2_b200 import _lib, synth

L = _lib.lib()
vp = C.c_void_p
L.orbm_knn2_mma_device.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp]
dev = torch.device("cuda", 0)
L.orbm_set_knn2_backend(1)  # orbm_knn2_device = the POPC kernel in this script


def run(fn, A, B):
    n = len(A)
    out = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)]
    rc = fn(vp(A.data_ptr()), n, vp(B.data_ptr()), len(B), vp(out[0].data_ptr()), vp(out[1].data_ptr()), vp(out[2].data_ptr()),
            vp(torch.cuda.current_stream().cuda_stream))
    if rc:
        print("rc", rc, L.orb_last_error()); sys.exit(1)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


ok = True
for (na, nb, seed) in [(128, 256, 0), (1, 1, 1), (129, 257, 2), (1000, 1000, 3), (2013, 2013, 4), (300, 5000, 5), (5000, 300, 6), (1013, 77, 7), (4000, 20000, 8)]:
    Bh = synth.descriptors(nb, seed)
    Ah = synth.descriptors(na, seed + 100, dup_from=Bh, max_flips=60)
    if nb > 10:
        Bh[nb // 2] = Bh[3]; Bh[nb - 1] = Bh[3]  # exact duplicates: first-minimum and tie handling
    A, B = torch.as_tensor(Ah).to(dev), torch.as_tensor(Bh).to(dev)
    want = run(L.orbm_knn2_device, A, B)
    got = run(L.orbm_knn2_mma_device, A, B)
    same = all(np.array_equal(g, w) for g, w in zip(got, want))
    ok &= same
    print(na, nb, "equal" if same else "DIFFERENT", [int((g != w).sum()) for g, w in zip(got, want)])
    if not same:
        bad = np.flatnonzero((got[0] != want[0]) | (got[1] != want[1]) | (got[2] != want[2]))[:5]
        for i in bad:
            print("  row", i, "got", [int(g[i]) for g in got], "want", [int(w[i]) for w in want])
if not ok:
    sys.exit(1)
for (na, nb) in [(1000, 1000), (20000, 20000), (100000, 100000), (200000, 200000)]:
    rng = np.random.default_rng(na)
    Ah = rng.integers(0, 256, (na, 32), dtype=np.uint8)
    Bh = rng.integers(0, 256, (nb, 32), dtype=np.uint8)
    A, B = torch.as_tensor(Ah).to(dev), torch.as_tensor(Bh).to(dev)
    for name, fn in (("popc", L.orbm_knn2_device), ("mma ", L.orbm_knn2_mma_device)):
        run(fn, A, B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = [torch.empty(na, dtype=torch.int32, device=dev) for _ in range(3)]
        reps = 5
        e0.record()
        for _ in range(reps):
            fn(vp(A.data_ptr()), na, vp(B.data_ptr()), nb, vp(out[0].data_ptr()), vp(out[1].data_ptr()), vp(out[2].data_ptr()),
               vp(torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(name, na, nb, "%.3f ms  %.1f Gcmp/s" % (ms, na * nb / ms / 1e6))
