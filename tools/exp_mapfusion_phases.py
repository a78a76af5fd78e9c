"""Per-phase timing of CrossMapMatcher.match under torchrun (diagnostic)."""
import ctypes as C
import os

import torch
import torch.distributed as dist

from multiagent_orb_slam2_b200 import _lib, mapfusion, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
rows = 200000
n_maps = max(2, world)
mine = [m for m in range(n_maps) if mapfusion.owner_of_map(m, world) == rank]
base = synth.descriptors(rows, 4242)
local = [torch.from_numpy(synth.descriptors_fast(rows, 5000 + m, base, 60)).to(dev) for m in mine]
cm = mapfusion.CrossMapMatcher(rows, 0.75)
cm.match(local)
dist.barrier()
L = _lib.lib()
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    sets, counts = mapfusion.exchange(local, rows)
    ev[1].record()
    pairs = mapfusion.plan_pairs(sets.shape[0], world, rank)
    p = torch.tensor(pairs, dtype=torch.int32, device=dev).reshape(-1, 2)
    n = len(pairs)
    out = [torch.empty((n, rows), dtype=torch.int32, device=dev) for _ in range(4)]
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ev[2].record()
    _lib.check(L.orbm_knn2_pairs_device(C.c_void_p(sets.data_ptr()), C.c_void_p(counts.data_ptr()), rows, C.c_void_p(p.data_ptr()), n,
                                        C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), C.c_void_p(out[2].data_ptr()), st))
    ev[3].record()
    torch.cuda.synchronize()
    print("rank", rank, "rep", rep, "pairs", pairs, "exchange %.2f  setup %.2f  knn2 %.2f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])), flush=True)
    dist.barrier()
dist.destroy_process_group()
