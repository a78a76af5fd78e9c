"""MapFusion cross-map candidate matching across GPUs (BASELINE config 5).

Every rank owns the keyframe descriptor set of one (or more) agent maps. One exchange step - an
NCCL all-gather of the descriptor sets over NVLink - makes every set resident on every GPU; the
directed map pairs (query map a, database map b != a) are then dealt round-robin to the ranks and
matched with the brute-force kNN-2 kernel (orbm_knn2_pairs_device). A query's (best, second, index)
is computed wholly on one rank in canonical database order, so the result is identical to the
single-GPU one. This is the SearchByBoW(KF,KF) inner loop (/root/reference/src/ORBmatcher.cc:566-603,
called from MapFusion::ComputeSim3 src/MapFusion.cc:275 and CovisibilityDiscovery :849) with the
vocabulary gate removed.

torch.distributed is plumbing only (process group, all-gather); all compute goes through the C ABI."""
import ctypes as C

import numpy as np

from .matcher import ORBmatcher


def directed_pairs(n_maps):
    """All (query map, database map) pairs, a != b, in canonical order."""
    return [(a, b) for a in range(n_maps) for b in range(n_maps) if a != b]


def plan_pairs(n_maps, world, rank):
    """Pairs handled by `rank`: round-robin over the canonical list. Disjoint over ranks, complete."""
    return directed_pairs(n_maps)[rank::world]


def owner_of_map(m, world):
    return m % world


def gather_layout(local_sets, rows_cap):
    """Packs this rank's sets into the fixed-size block that the all-gather exchanges:
    (sets_per_rank, rows_cap, 32) uint8 + (sets_per_rank,) int32 counts."""
    import torch
    k = len(local_sets)
    dev = local_sets[0].device if k else "cpu"
    block = torch.zeros((k, rows_cap, 32), dtype=torch.uint8, device=dev)
    counts = torch.zeros(k, dtype=torch.int32, device=dev)
    for i, s in enumerate(local_sets):
        n = s.shape[0]
        if n > rows_cap:
            raise ValueError("descriptor set of %d rows exceeds rows_cap=%d" % (n, rows_cap))
        block[i, :n] = s
        counts[i] = n
    return block, counts


def exchange(local_sets, rows_cap, group=None):
    """All-gather of the descriptor sets. Map m lives on rank m % world at local slot m // world
    (every rank holds the same number of sets); the gathered array is re-ordered to map order.
    Returns (sets [M, rows_cap, 32], counts [M])."""
    import torch
    import torch.distributed as dist
    block, counts = gather_layout(local_sets, rows_cap)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return block, counts
    world = dist.get_world_size(group)
    k = block.shape[0]
    all_sets = torch.empty((world * k, rows_cap, 32), dtype=torch.uint8, device=block.device)
    all_counts = torch.empty(world * k, dtype=torch.int32, device=block.device)
    dist.all_gather_into_tensor(all_sets, block, group=group)  # rank-major concatenation
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    if k == 1:
        return all_sets, all_counts  # one map per rank: rank order is map order
    # (rank, slot) -> map index slot*world + rank
    sets = all_sets.view(world, k, rows_cap, 32).permute(1, 0, 2, 3).reshape(world * k, rows_cap, 32)
    cnts = all_counts.view(world, k).permute(1, 0).reshape(world * k)
    return sets.contiguous(), cnts.contiguous()


class CrossMapMatcher:
    """rows_cap: capacity (rows) of one map's descriptor set. match_fn is injectable for the CPU
    (gloo) tests of the host logic; the default is the CUDA kernel and needs a GPU."""

    def __init__(self, rows_cap, nnratio=0.75, th=ORBmatcher.TH_LOW, group=None, match_fn=None):
        self.rows_cap, self.nnratio, self.th, self.group = int(rows_cap), float(np.float32(nnratio)), int(th), group
        self.match_fn = match_fn or self._match_cuda

    def _match_cuda(self, sets, counts, pairs):
        import torch
        from . import _lib
        L = _lib.lib()
        dev = sets.device
        if dev.type != "cuda":
            raise _lib.OrbError(_lib.ORB_ECUDA, "cross-map matching needs CUDA tensors (no CPU fallback)")
        key = (tuple(pairs), str(dev))
        if getattr(self, "_pairs_key", None) != key:  # the pair list of a rank is static: upload it once
            self._pairs_dev = torch.tensor(pairs, dtype=torch.int32, device=dev).reshape(-1, 2)
            self._pairs_key = key
        p = self._pairs_dev
        n = len(pairs)
        out = [torch.empty((n, self.rows_cap), dtype=torch.int32, device=dev) for _ in range(4)]
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.orbm_knn2_pairs_device(C.c_void_p(sets.data_ptr()), C.c_void_p(counts.data_ptr()), self.rows_cap, C.c_void_p(p.data_ptr()), n,
                                            C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), C.c_void_p(out[2].data_ptr()), st))
        _lib.check(L.orbm_ratio_filter_device(C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), C.c_void_p(out[2].data_ptr()),
                                              n * self.rows_cap, self.th, 0, self.nnratio, C.c_void_p(out[3].data_ptr()), st))
        return out  # idx, best, second, match : each (pairs, rows_cap)

    def match(self, local_sets):
        """local_sets: list of (n_i, 32) uint8 tensors owned by this rank. Returns
        {(a, b): (idx, best, second, match)} for the pairs dealt to this rank (rows beyond the query
        map's count are undefined)."""
        import torch.distributed as dist
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if world > 1 else 0
        sets, counts = exchange(local_sets, self.rows_cap, self.group)
        pairs = plan_pairs(sets.shape[0], world, rank)
        if not pairs:
            return {}, counts
        idx, b1, b2, match = self.match_fn(sets, counts, pairs)
        return {pq: (idx[i], b1[i], b2[i], match[i]) for i, pq in enumerate(pairs)}, counts
