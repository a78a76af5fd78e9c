"""Multi-GPU MapFusion check (run under torchrun on >= 2 GPUs; not a pytest file):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_mapfusion_check.py
Every rank owns one map; the NCCL exchange + pair-sharded matching must reproduce, pair by pair, what a single GPU
computes on the gathered sets, and agree with the oracle on sampled rows."""
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
os.environ.pop("NCCL_DEBUG", None)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows = 20000
base = synth.descriptors(rows, 99)
maps = [synth.descriptors_fast(rows - 13 * m, 200 + m, base, 60) for m in range(world)]
cm = mapfusion.CrossMapMatcher(rows, 0.75)
res, counts = cm.match([torch.from_numpy(maps[rank]).cuda()])
torch.cuda.synchronize()
assert counts.cpu().tolist() == [len(m) for m in maps]
rng = np.random.default_rng(rank)
for (a, b), (idx, d1, d2, match) in res.items():
    n = len(maps[a])
    sample = rng.choice(n, 64, replace=False)
    oi, o1, o2 = O.knn2(maps[a][sample], maps[b])
    assert np.array_equal(idx.cpu().numpy()[sample], oi) and np.array_equal(d1.cpu().numpy()[sample], o1)
    assert np.array_equal(d2.cpu().numpy()[sample], o2)
pairs = torch.tensor([len(res)], device="cuda")
dist.all_reduce(pairs)
assert int(pairs.item()) == world * (world - 1)
if rank == 0:
    print("dist_mapfusion_check ok: %d ranks, %d directed pairs, sampled rows bit-exact vs oracle" % (world, int(pairs.item())))
dist.destroy_process_group()
