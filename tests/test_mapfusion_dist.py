"""Multi-GPU MapFusion path (csrc/xmap.cu, multiagent_orb_slam2_b200/mapfusion.py).
CPU: the (pair, query tile) split of orbm_xmap_plan is a balanced partition for every maps / ranks combination incl.
fewer maps than ranks, and a world_size-2 gloo run in which every rank computes ITS chunks with the oracle reassembles to
the single-process result. GPU: one device (several maps on one GPU), several devices driven by one process
(orbm_xmap_attach_local) incl. 2 maps split over all GPUs, and the multi-process NCCL-free path under torch.distributed.run
(CUDA IPC windows) - all against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from multiagent_orb_slam2_b200 import mapfusion, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS = 700


def make_maps(n_maps, rows=ROWS, step=37):
    base = synth.descriptors(rows, 77)
    return [synth.descriptors(rows - step * m, 100 + m, dup_from=base, max_flips=50) for m in range(n_maps)]


def tiles_of(rows):
    return (rows + 127) // 128


@pytest.mark.parametrize("rows", [[700, 663, 626], [200000, 200000], [1, 129, 128, 0], [5000] * 8, [300, 70000]])
def test_chunk_plan_is_a_balanced_partition(rows):
    n = len(rows)
    for world in (1, 2, 3, 4, 8, 16):
        seen = {}
        loads = []
        for r in range(world):
            chunks = mapfusion.plan_chunks(rows, world, r)
            load = 0
            for a, b, t0, t1 in chunks:
                assert a != b and 0 <= t0 < t1 <= tiles_of(rows[a])
                for t in range(t0, t1):
                    assert (a, b, t) not in seen
                    seen[(a, b, t)] = r
                load += t1 - t0
            loads.append(load)
        want = {(a, b, t) for a, b in mapfusion.directed_pairs(n) for t in range(tiles_of(rows[a]))}
        assert set(seen) == want
        assert max(loads) - min(loads) <= 1
    # as many equal maps as ranks: every rank matches exactly its own map against the others (queries stay local)
    if len(set(rows)) == 1:
        for r in range(n):
            assert {c[0] for c in mapfusion.plan_chunks(rows, n, r)} == {r}


def test_two_maps_keep_eight_ranks_busy():
    for r in range(8):
        chunks = mapfusion.plan_chunks([200000, 200000], 8, r)
        assert sum(t1 - t0 for _, _, t0, t1 in chunks) in (390, 391)   # 2 x 1563 tiles over 8 ranks


def _worker(rank, world, port, n_maps, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        maps = make_maps(n_maps, 300, 11)
        local_rows = [len(maps[m]) for m in range(n_maps) if mapfusion.owner_of_map(m, world) == rank]
        rows = mapfusion.gather_rows(local_rows, n_maps)           # the set-up exchange, over gloo here
        out = []
        for a, b, t0, t1 in mapfusion.plan_chunks(rows, world, rank):
            r0, r1 = t0 * 128, min(t1 * 128, rows[a])
            out.append((a, b, r0, O.knn2(maps[a][r0:r1], maps[b])))   # this rank's share, by the oracle
        q.put((rank, rows, out))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure instead of letting the parent time out
        q.put((rank, repr(e), None))
        raise


@pytest.mark.parametrize("n_maps", [2, 3])
def test_two_rank_gloo_chunks_reassemble_to_the_single_process_result(n_maps):
    world, port = 2, 29500 + np.random.default_rng().integers(0, 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, int(port), n_maps, q)) for r in range(world)]
    for p in procs:
        p.start()
    maps = make_maps(n_maps, 300, 11)
    got = {pq: [np.full(len(maps[pq[0]]), -7, np.int32) for _ in range(3)] for pq in mapfusion.directed_pairs(n_maps)}
    for _ in range(world):
        rank, rows, out = q.get(timeout=120)
        assert out is not None, rows
        assert rows == [len(m) for m in maps]
        for a, b, r0, (idx, d1, d2) in out:
            for dst, src in zip(got[(a, b)], (idx, d1, d2)):
                assert np.all(dst[r0:r0 + len(src)] == -7)      # no row is produced twice
                dst[r0:r0 + len(src)] = src
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for (a, b), (idx, d1, d2) in got.items():
        oi, o1, o2 = O.knn2(maps[a], maps[b])
        assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2)


def check_against_oracle(res, maps, sample=None, seed=0):
    rng = np.random.default_rng(seed)
    for (a, b), (idx, d1, d2, match) in res.items():
        n = len(maps[a])
        rows = np.arange(n) if sample is None or n <= sample else np.sort(rng.choice(n, sample, replace=False))
        oi, o1, o2 = O.knn2(maps[a][rows], maps[b])
        gi, g1, g2, gm = (t.cpu().numpy()[rows] for t in (idx, d1, d2, match))
        assert np.array_equal(gi, oi) and np.array_equal(g1, o1) and np.array_equal(g2, o2), (a, b)
        ref = np.where((o1 < 50) & (o1.astype(np.float32) < np.float32(0.75) * o2.astype(np.float32)), oi, -1)
        assert np.array_equal(gm, ref), (a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("n_maps", [2, 3])
def test_single_gpu_cross_map_matching_matches_oracle(n_maps):
    maps = make_maps(n_maps)
    cm = mapfusion.CrossMapMatcher(ROWS, n_maps, 0.75)
    d = [torch.from_numpy(m).cuda() for m in maps]
    for step in range(4):   # several steps: both halves of the double-buffered window, flags advance
        res, rows = cm.match(d)
        torch.cuda.synchronize()
        assert rows == [len(m) for m in maps] and sorted(res) == sorted(mapfusion.directed_pairs(n_maps))
        check_against_oracle(res, maps)
        # new content in the same buffers for the next step
        maps = [np.roll(m, 5 * (step + 1), axis=0).copy() for m in maps]
        for t, m in zip(d, maps):
            t.copy_(torch.from_numpy(m))
    cm.close()


@pytest.mark.gpu
def test_single_gpu_edge_cases_empty_and_ragged_maps():
    maps = [synth.descriptors(n, 300 + i) if n else np.zeros((0, 32), np.uint8) for i, n in enumerate([129, 0, 1, 256])]
    cm = mapfusion.CrossMapMatcher(256, 4, 0.75)
    res, rows = cm.match([torch.from_numpy(m).cuda() for m in maps])
    torch.cuda.synchronize()
    for (a, b), (idx, d1, d2, match) in res.items():
        n = len(maps[a])
        if n == 0:
            continue
        if len(maps[b]) == 0:   # no candidates: the initial state of the reference loop
            assert np.all(idx.cpu().numpy() == -1) and np.all(d1.cpu().numpy() == 256) and np.all(d2.cpu().numpy() == 256)
            continue
        oi, o1, o2 = O.knn2(maps[a], maps[b])
        assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(d1.cpu().numpy(), o1) and np.array_equal(d2.cpu().numpy(), o2)
    cm.close()


def _local_world(n_maps, world, rows_cap, maps, steps=2):
    """`world` contexts in THIS process, one per GPU (orbm_xmap_attach_local): what a single-process multi-agent server does."""
    from multiagent_orb_slam2_b200 import _lib
    L = _lib.lib()
    ctxs = (C.c_void_p * world)()
    for r in range(world):
        h = C.c_void_p()
        _lib.check(L.orbm_xmap_create(r, r, world, n_maps, rows_cap, C.byref(h)))
        ctxs[r] = h
    _lib.check(L.orbm_xmap_attach_local(ctxs, world))
    rows = np.array([len(m) for m in maps], np.int32)
    d = {m: torch.from_numpy(maps[m]).to("cuda:%d" % (m % world)) for m in range(n_maps)}
    streams = [torch.cuda.Stream("cuda:%d" % r) for r in range(world)]
    for _ in range(steps):
        for r in range(world):
            own = [d[m] for m in range(r, n_maps, world)]
            ptrs = (C.c_void_p * max(1, len(own)))(*[C.c_void_p(t.data_ptr()) for t in own])
            _lib.check(L.orbm_knn2_allgather(ctxs[r], ptrs, rows.ctypes.data_as(C.c_void_p), C.c_void_p(streams[r].cuda_stream)))
        for s in streams:
            s.synchronize()
    out = {}
    for a, b in mapfusion.directed_pairs(n_maps):
        r = a % world
        p = [C.c_void_p() for _ in range(3)]
        _lib.check(L.orbm_xmap_result(ctxs[r], a, b, *[C.byref(x) for x in p]))
        out[(a, b)] = [mapfusion._device_view(x.value, len(maps[a]), torch.device("cuda", r)).cpu().numpy() for x in p]
    for r in range(world):
        L.orbm_xmap_destroy(ctxs[r])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("n_maps", [2, 5])
def test_several_gpus_in_one_process_incl_query_row_split(n_maps):
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    maps = make_maps(n_maps, 3000, 101)
    for world in sorted({2, min(ndev, 4), ndev}):
        got = _local_world(n_maps, world, 3000, maps)
        for (a, b), (idx, d1, d2) in got.items():
            oi, o1, o2 = O.knn2(maps[a], maps[b])
            assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2), (world, a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("n_maps_per_world", ["world", "2"])
def test_multi_process_windows_over_cuda_ipc(n_maps_per_world):
    """One process per GPU under torch.distributed.run (the way bench.py --gpus N is launched): window handles travel
    once over the process group, the data path is peer loads / stores only. tests/dist_mapfusion_check.py checks sampled
    rows of every pair against the oracle on every rank."""
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    port = 29500 + int(np.random.default_rng().integers(0, 2000))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ndev), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_mapfusion_check.py"), n_maps_per_world]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "dist_mapfusion_check ok" in out.stdout


def build_xmap_driver(tmp):
    exe = os.path.join(tmp, "xmap_test")
    lib = os.path.join(ROOT, "multiagent_orb_slam2_b200", "lib")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "cpp", "xmap_test.cc"), "-o", exe, "-L" + lib, "-lorb_b200", "-Wl,-rpath," + lib,
                           "-L/usr/local/cuda/lib64", "-lcudart"])
    return exe


def test_cpp_xmap_driver_compiles(tmp_path):
    assert os.path.exists(build_xmap_driver(str(tmp_path)))


@pytest.mark.gpu
@pytest.mark.parametrize("n_maps,rows", [(2, 5000), (3, 1300)])
def test_cpp_driver_cross_map_matching(tmp_path, n_maps, rows):
    """orbm_knn2_allgather driven from C++ in one process over every visible GPU (1 on the test box; with --gpus N the
    query rows of the maps are split over N GPUs)."""
    exe = build_xmap_driver(str(tmp_path))
    out = subprocess.run([exe, str(n_maps), str(rows)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "xmap ok" in out.stdout
