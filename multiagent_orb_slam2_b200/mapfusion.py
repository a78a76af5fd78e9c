"""MapFusion cross-map candidate matching across GPUs (BASELINE config 5), host side.

Every rank (GPU) owns the keyframe descriptor set of one or more agent maps: map m lives on rank m % world. One call of
orbm_knn2_allgather per rank does the exchange AND the matching (csrc/xmap.cu): ranks read the packed descriptor sets they
need straight out of their owners' device windows over NVLink inside the kernel that expands them into the tensor-core
operands, the (query map, database map) pairs are cut into 128-row query tiles that are dealt evenly to the ranks (so
2 maps keep 8 GPUs busy), and results land in the window of the query map's owner. No collective library call is on the
data path; a query's (best, second, index) is computed over the whole database in canonical order, so it equals the
single-GPU result. This is the SearchByBoW(KF,KF) inner loop (/root/reference/src/ORBmatcher.cc:566-603, called from
MapFusion::ComputeSim3 src/MapFusion.cc:275 and CovisibilityDiscovery :849) with the vocabulary gate removed.

torch.distributed is set-up plumbing only: it carries the 64-byte window handles (once) and the row counts (when they
change) between the processes. All compute goes through the C ABI."""
import ctypes as C

import numpy as np

from .matcher import ORBmatcher

TILE = 128


def directed_pairs(n_maps):
    """All (query map, database map) pairs, a != b, in canonical order."""
    return [(a, b) for a in range(n_maps) for b in range(n_maps) if a != b]


def owner_of_map(m, world):
    return m % world


def plan_chunks(rows_per_map, world, rank):
    """The (a, b, tile0, tile1) chunks of the flattened (pair, 128-row query tile) list that `rank` handles: the C ABI's
    own split (orbm_xmap_plan, host only - no device needed)."""
    from . import _lib
    L = _lib.lib()
    rows = np.ascontiguousarray(rows_per_map, np.int32)
    n = C.c_int(0)
    cap = 2 * len(rows) * len(rows) + 4
    out = np.zeros((cap, 4), np.int32)
    _lib.check(L.orbm_xmap_plan(len(rows), rows.ctypes.data_as(C.c_void_p), world, rank, out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    return [tuple(int(v) for v in r) for r in out[:n.value]]


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def gather_rows(local_rows, n_maps, group=None, device=None):
    """rows_per_map (list of n_maps ints) from the counts of the maps each rank owns (set-up plumbing)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group) if dist else 1
    rank = dist.get_rank(group) if dist else 0
    slots = (n_maps + world - 1) // world
    mine = torch.full((slots,), -1, dtype=torch.int32)
    mine[:len(local_rows)] = torch.tensor(list(local_rows), dtype=torch.int32)
    if world == 1:
        allr = mine.view(1, slots)
    else:
        t = mine.to(device) if device is not None else mine
        out = torch.empty(world * slots, dtype=torch.int32, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
        allr = out.cpu().view(world, slots)
    rows = [int(allr[m % world, m // world]) for m in range(n_maps)]
    if any(r < 0 for r in rows):
        raise ValueError("rank %d: some map has no owner-provided row count: %r" % (rank, rows))
    return rows


class CrossMapMatcher:
    """n_maps maps of up to rows_cap descriptors over the ranks of `group` (one process per GPU), or over one GPU when
    torch.distributed is not initialised. match(local_sets) runs one step; results are returned for the pairs whose
    query map this rank owns."""

    def __init__(self, rows_cap, n_maps, nnratio=0.75, th=ORBmatcher.TH_LOW, group=None, device=None):
        import torch
        from . import _lib
        self.L = _lib.lib()
        self.rows_cap, self.n_maps = int(rows_cap), int(n_maps)
        self.nnratio, self.th, self.group = float(np.float32(nnratio)), int(th), group
        dist = _dist()
        self.world = dist.get_world_size(group) if dist else 1
        self.rank = dist.get_rank(group) if dist else 0
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        h = C.c_void_p()
        _lib.check(self.L.orbm_xmap_create(self.device, self.rank, self.world, self.n_maps, self.rows_cap, C.byref(h)))
        self._h = h
        if self.world > 1:
            mine = (C.c_ubyte * 64)()
            _lib.check(self.L.orbm_xmap_ipc_handle(h, mine))
            t = torch.tensor(list(mine), dtype=torch.uint8, device=self.dev)
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(allh, t, group=group)
            blob = allh.cpu().numpy().tobytes()
            _lib.check(self.L.orbm_xmap_attach_ipc(h, blob))
        self.owned = [m for m in range(self.n_maps) if owner_of_map(m, self.world) == self.rank]
        self._rows_key, self._rows = None, None
        self._match = {}

    def close(self):
        if getattr(self, "_h", None):
            self.L.orbm_xmap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rows_per_map(self, local_sets):
        key = tuple(int(s.shape[0]) for s in local_sets)
        if key != self._rows_key:
            self._rows = gather_rows(key, self.n_maps, self.group, self.dev)
            self._rows_key = key
        return self._rows

    def step(self, local_sets, rows_per_map=None):
        """Enqueues one exchange + match step on torch's current stream. local_sets: the maps this rank owns, in map
        order, as (n, 32) uint8 CUDA tensors."""
        import torch
        from . import _lib
        if len(local_sets) != len(self.owned):
            raise ValueError("rank %d owns %d maps, got %d sets" % (self.rank, len(self.owned), len(local_sets)))
        rows = list(rows_per_map) if rows_per_map is not None else self.rows_per_map(local_sets)
        ptrs = (C.c_void_p * max(1, len(local_sets)))(*[C.c_void_p(s.data_ptr()) for s in local_sets])
        r = np.ascontiguousarray(rows, np.int32)
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(self.L.orbm_knn2_allgather(self._h, ptrs, r.ctypes.data_as(C.c_void_p), st))
        return rows

    def result(self, a, b, n_rows):
        """(idx, best, second) of pair (a, b) as int32 CUDA tensors viewing the window (valid once the stream is done)."""
        import torch
        from . import _lib
        p = [C.c_void_p() for _ in range(3)]
        _lib.check(self.L.orbm_xmap_result(self._h, a, b, *[C.byref(x) for x in p]))
        return [_device_view(x.value, n_rows, self.dev) for x in p]

    def match(self, local_sets, rows_per_map=None):
        """One step + the SearchByBoW(KF,KF) acceptance test (src/ORBmatcher.cc:600-603). Returns
        ({(a, b): (idx, best, second, match)} for the pairs whose query map this rank owns, rows_per_map)."""
        import torch
        from . import _lib
        rows = self.step(local_sets, rows_per_map)
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        res = {}
        for a in self.owned:
            for b in range(self.n_maps):
                if b == a:
                    continue
                idx, b1, b2 = self.result(a, b, rows[a])
                m = self._match.get((a, b))
                if m is None or m.shape[0] < rows[a]:
                    m = self._match[(a, b)] = torch.empty(self.rows_cap, dtype=torch.int32, device=self.dev)
                if rows[a]:
                    _lib.check(self.L.orbm_ratio_filter_device(C.c_void_p(idx.data_ptr()), C.c_void_p(b1.data_ptr()), C.c_void_p(b2.data_ptr()),
                                                               rows[a], self.th, 0, self.nnratio, C.c_void_p(m.data_ptr()), st))
                res[(a, b)] = (idx, b1, b2, m[:rows[a]])
        return res, rows


def _device_view(ptr, n, dev):
    """An int32 CUDA tensor of n elements over existing device memory (no copy, no ownership)."""
    import torch

    class _Ext:
        pass
    e = _Ext()
    e.__cuda_array_interface__ = {"shape": (max(n, 1),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
    with torch.cuda.device(dev):
        t = torch.as_tensor(e, device=dev)
    return t[:n]
