"""CTA-pair tensor-core kNN-2 (backend 3, tcgen05 cta_group::2) against the one-CTA kernel (backend 2) and the POPC kernel
(backend 1, oracle-verified): results and speed. Run each size in a bounded loop; any device fault ends the script."""
import ctypes as C
import sys

import numpy as np
import torch

from multiagent_orb_slam2_b200 import _lib, synth

L = _lib.lib()
vp = C.c_void_p
dev = torch.device("cuda", 0)


def run(backend, A, B, reps=1):
    _lib.check(L.orbm_set_knn2_backend(backend))
    n = len(A)
    out = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)]
    st = vp(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        rc = L.orbm_knn2_device(vp(A.data_ptr()), n, vp(B.data_ptr()), len(B), vp(out[0].data_ptr()), vp(out[1].data_ptr()), vp(out[2].data_ptr()), st)
        if rc:
            print("rc", rc, L.orb_last_error()); sys.exit(1)
    e1.record()
    torch.cuda.synchronize()
    _lib.check(L.orbm_set_knn2_backend(0))
    return [o.cpu().numpy() for o in out], e0.elapsed_time(e1) / reps


ok = True
for (na, nb, seed) in [(256, 128, 0), (128, 256, 1), (1, 1, 2), (129, 257, 3), (257, 129, 4), (1000, 1000, 5), (2013, 2013, 6), (300, 5000, 7),
                       (5000, 300, 8), (1013, 77, 9), (4000, 20000, 10), (385, 9000, 11), (20000, 4097, 12)]:
    Bh = synth.descriptors(nb, seed)
    Ah = synth.descriptors(na, seed + 100, dup_from=Bh, max_flips=60)
    if nb > 10:
        Bh[nb // 2] = Bh[3]; Bh[nb - 1] = Bh[3]
    A, B = torch.as_tensor(Ah).to(dev), torch.as_tensor(Bh).to(dev)
    want, _ = run(1, A, B)
    got, _ = run(3, A, B)
    same = all(np.array_equal(g, w) for g, w in zip(got, want))
    ok &= same
    print(na, nb, "equal" if same else "DIFFERENT", [int((g != w).sum()) for g, w in zip(got, want)], flush=True)
    if not same:
        bad = np.flatnonzero((got[0] != want[0]) | (got[1] != want[1]) | (got[2] != want[2]))[:5]
        for i in bad:
            print("  row", i, "got", [int(g[i]) for g in got], "want", [int(w[i]) for w in want])
if not ok:
    sys.exit(1)
for (na, nb) in [(1000, 1000), (2000, 4096), (8000, 8000), (20000, 20000), (100000, 100000), (200000, 200000), (25000, 200000)]:
    rng = np.random.default_rng(na)
    A = torch.as_tensor(rng.integers(0, 256, (na, 32), dtype=np.uint8)).to(dev)
    B = torch.as_tensor(rng.integers(0, 256, (nb, 32), dtype=np.uint8)).to(dev)
    res = {}
    for name, backend in (("one-cta", 2), ("pair   ", 3)):
        run(backend, A, B)
        res[name], ms = run(backend, A, B, reps=5)
        print(name, na, nb, "%.3f ms  %.1f Gcmp/s" % (ms, na * nb / ms / 1e6), flush=True)
    print("   identical:", all(np.array_equal(a, b) for a, b in zip(res["one-cta"], res["pair   "])))
