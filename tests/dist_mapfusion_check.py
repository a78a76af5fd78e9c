"""Multi-GPU MapFusion check, run under torchrun on >= 2 GPUs by tests/test_mapfusion_dist.py (or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_mapfusion_check.py [world|2]
"world": one map per rank; "2": two maps on ranks 0 and 1, their query rows split over ALL ranks. Every pair's sampled rows
must equal the oracle's, on the rank that owns the query map."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_lib as O  # noqa: E402
from multiagent_orb_slam2_b200 import mapfusion, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_maps = world if (len(sys.argv) < 2 or sys.argv[1] == "world") else int(sys.argv[1])
rows = 20000
base = synth.descriptors(rows, 99)
maps = [synth.descriptors_fast(rows - 13 * m, 200 + m, base, 60) for m in range(n_maps)]
cm = mapfusion.CrossMapMatcher(rows, n_maps, 0.75)
mine = [torch.from_numpy(maps[m]).cuda() for m in cm.owned]
for step in range(3):
    res, counts = cm.match(mine)
    torch.cuda.synchronize()
    assert counts == [len(m) for m in maps]
    rng = np.random.default_rng(rank + 10 * step)
    for (a, b), (idx, d1, d2, match) in res.items():
        n = len(maps[a])
        sample = np.sort(rng.choice(n, 96, replace=False))
        sample[0], sample[-1] = 0, n - 1
        oi, o1, o2 = O.knn2(maps[a][sample], maps[b])
        assert np.array_equal(idx.cpu().numpy()[sample], oi) and np.array_equal(d1.cpu().numpy()[sample], o1), (rank, a, b)
        assert np.array_equal(d2.cpu().numpy()[sample], o2), (rank, a, b)
    # next step: rotated content in the same device buffers
    maps = [np.roll(m, 17, axis=0).copy() for m in maps]
    for t, m in zip(mine, cm.owned):
        t.copy_(torch.from_numpy(maps[m]))
    dist.barrier()   # test hygiene only: every rank rotates before anyone publishes the next step
pairs = torch.tensor([len(res)], device="cuda")
dist.all_reduce(pairs)
assert int(pairs.item()) == n_maps * (n_maps - 1)
if rank == 0:
    print("dist_mapfusion_check ok: %d ranks, %d maps, %d directed pairs, sampled rows bit-exact vs oracle" % (world, n_maps, int(pairs.item())))
cm.close()
dist.destroy_process_group()
