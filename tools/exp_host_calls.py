"""Wall-clock cost of the host-buffer matcher entry points (what the C++ facade calls once per search)."""
import time

import numpy as np

from multiagent_orb_slam2_b200 import ORBmatcher, synth

m = ORBmatcher(0.75)
B = synth.descriptors(1000, 1)
A = synth.descriptors_fast(1000, 2, B, 50)
rng = np.random.default_rng(0)
offsets = np.arange(0, 1001 * 20, 20, dtype=np.int32)
cands = rng.integers(0, 1000, 1000 * 20).astype(np.int32)


def t(fn, reps=200):
    for _ in range(10):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


print("knn2 1000x1000 host call      : %.1f us" % t(lambda: m.knn2(A, B)))
print("knn2_lists 1000 x 20 host call: %.1f us" % t(lambda: m.knn2_lists(A, B, offsets, cands)))
print("distance_matrix 30x30         : %.1f us" % t(lambda: m.distance_matrix(A[:30], B[:30])))
