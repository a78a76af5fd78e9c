"""Host logic of the multi-GPU MapFusion path (exchange layout, pair planning, result placement) on
CPU with world_size-2 gloo; the matcher is the oracle here (injected), the CUDA kernel on the GPU box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from multiagent_orb_slam2_b200 import mapfusion, synth
from multiagent_orb_slam2_b200.matcher import ORBmatcher

ROWS = 160


def make_maps(n_maps):
    base = synth.descriptors(ROWS, 77)
    return [synth.descriptors(ROWS - 7 * m, 100 + m, dup_from=base, max_flips=50) for m in range(n_maps)]


def oracle_match_fn(sets, counts, pairs):
    m = ORBmatcher.__new__(ORBmatcher)
    m.mfNNratio, m.TH_LOW = float(np.float32(0.75)), 50
    out = [torch.zeros((len(pairs), ROWS), dtype=torch.int32) for _ in range(4)]
    for i, (a, b) in enumerate(pairs):
        na, nb = int(counts[a]), int(counts[b])
        idx, d1, d2 = O.knn2(sets[a, :na].numpy(), sets[b, :nb].numpy())
        acc = np.where((d1 < 50) & (d1.astype(np.float32) < np.float32(0.75) * d2.astype(np.float32)), idx, -1)
        for t, v in zip(out, (idx, d1, d2, acc)):
            t[i, :na] = torch.from_numpy(v.astype(np.int32))
    return out


def test_pair_plan_is_a_partition():
    for n_maps in (2, 3, 8):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                got += mapfusion.plan_pairs(n_maps, world, r)
            assert sorted(got) == sorted(mapfusion.directed_pairs(n_maps)) and len(set(got)) == len(got)


def _worker(rank, world, port, n_maps, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        maps = make_maps(n_maps)
        local = [torch.from_numpy(maps[m]) for m in range(n_maps) if mapfusion.owner_of_map(m, world) == rank]
        cm = mapfusion.CrossMapMatcher(ROWS, 0.75, match_fn=oracle_match_fn)
        res, counts = cm.match(local)
        q.put((rank, {k: [t.numpy() for t in v] for k, v in res.items()}, counts.numpy()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure instead of letting the parent time out
        q.put((rank, repr(e), None))
        raise


@pytest.mark.parametrize("n_maps", [2, 4])
def test_two_rank_gloo_equals_single_process(n_maps):
    world, port = 2, 29500 + np.random.default_rng().integers(0, 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, int(port), n_maps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    counts = None
    for _ in range(world):
        rank, res, cnt = q.get(timeout=120)
        assert isinstance(res, dict), res
        assert not (set(res) & set(got))
        got.update(res)
        counts = cnt
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    maps = make_maps(n_maps)
    assert counts.tolist() == [len(m) for m in maps]
    assert sorted(got) == sorted(mapfusion.directed_pairs(n_maps))
    for (a, b), (idx, d1, d2, match) in got.items():
        oi, o1, o2 = O.knn2(maps[a], maps[b])
        n = len(maps[a])
        assert np.array_equal(idx[:n], oi) and np.array_equal(d1[:n], o1) and np.array_equal(d2[:n], o2)
        assert (match[:n] >= 0).sum() > 10


@pytest.mark.gpu
def test_single_gpu_cross_map_matching_matches_oracle():
    maps = make_maps(3)
    cm = mapfusion.CrossMapMatcher(ROWS, 0.75)
    res, counts = cm.match([torch.from_numpy(m).cuda() for m in maps])
    torch.cuda.synchronize()
    assert sorted(res) == sorted(mapfusion.directed_pairs(3))
    for (a, b), (idx, d1, d2, match) in res.items():
        oi, o1, o2 = O.knn2(maps[a], maps[b])
        n = len(maps[a])
        assert np.array_equal(idx[:n].cpu().numpy(), oi) and np.array_equal(d1[:n].cpu().numpy(), o1) and np.array_equal(d2[:n].cpu().numpy(), o2)
        ref = np.where((o1 < 50) & (o1.astype(np.float32) < np.float32(0.75) * o2.astype(np.float32)), oi, -1)
        assert np.array_equal(match[:n].cpu().numpy(), ref)
