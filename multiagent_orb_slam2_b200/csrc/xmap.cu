// Cross-map descriptor matching across the GPUs of one node WITHOUT a collective library call on the data path
// (BASELINE config 5; SURVEY.md sections 5 and 8e: "MapFusion cross-agent matching all-gathers keyframe descriptor sets").
//
// Reference call sites: MapFusion::ComputeSim3 / CovisibilityDiscovery run ORBmatcher::SearchByBoW(KF, KF)
// (/root/reference/src/MapFusion.cc:275, 849 -> src/ORBmatcher.cc:524-657) between keyframes of different agents' maps; the
// reference shares pointers inside one process. Here every map's descriptor set lives on its agent's GPU, and
// orbm_knn2_allgather is the exchange AND the matching:
//
//   * every rank owns a WINDOW (one cudaMalloc, exported by cudaIpcGetMemHandle or, for several GPUs driven by one process,
//     peer-mapped directly): a control block, the packed descriptor sets of the maps it owns (double buffered by step
//     parity) and the result arrays of the pairs whose QUERY map it owns;
//   * publishing a set = copy into the window + a release-store of the step number (ready flag);
//   * the all-gather is fused into the operand expansion: the kernel that turns 32-byte descriptors into the +-1 / +-64 int8
//     operands of the tensor-core matcher (hamming_mma.cu) reads the set straight out of the OWNER's window over NVLink
//     (plain loads through the peer mapping, after an acquire-wait on the owner's ready flag) and writes the expanded
//     operand into local HBM, where the tcgen05 kernel streams it many times; 6.4 MB cross the link per 200 k-row map and pair,
//     nothing is staged in between and no rank waits for sets it does not need;
//   * the directed (query map, database map) pairs are cut into 128-row query tiles and the flattened tile list is dealt
//     evenly to the ranks: with as many maps as GPUs every rank matches its own map against all others, with FEWER maps than
//     GPUs the query rows of a map are split over several ranks, whose results go straight into the
//     owner's window by peer stores; a query's (best, second, index) is always computed over the whole database in canonical
//     order (by one CTA, or by 2 - 4 CTAs over consecutive candidate ranges whose partial results are folded in order), so the
//     result does not depend on the split;
//   * completion = a release-store of the step number into every owner's done[rank] flag; an owner's call ends with an
//     acquire-wait for all contributors, and a publish of step e waits until every rank is done with step e-2 (the step that
//     used the same half of the double buffer).
// Waits are bounded spins that trap instead of hanging the GPU when a peer never arrives.
#include <cuda.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace orb {
int mma_encode_operand_map(CUtensorMap* out, const void* base, long long rows, bool query_side);
int mma_expand_rows(const uint8_t* packed, int n_rows, bool query_side, uint8_t* out, cudaStream_t st);
int mma_encode_operand_map_half(CUtensorMap* out, const void* base, long long rows);
size_t mma_partial_ints(size_t rows);
int mma_launch_preexpanded(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_bh, int na, int nb, int* d_idx, int* d_b1,
                           int* d_b2, int* part, cudaStream_t st);
int mma_preload_kernels();

constexpr int kXmapMaxWorld = 16, kXmapMaxSlots = 8, kXmapCtrlBytes = 4096, kXmapTile = 128;
struct XmapCtrl {
    uint32_t ready[kXmapMaxSlots];   // step number of the set last published into slot s (half = step & 1)
    uint32_t done[kXmapMaxWorld];    // done[r]: last step rank r has finished (its reads of this window and its result stores)
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// lane i waits until flags[i] >= target (i < n). ~4 s of polling, then trap: a missing peer must not hang the box.
__global__ void xmap_wait_kernel(const uint32_t* const* flags, int n, uint32_t target) {
    const int i = threadIdx.x;
    if (i >= n) return;
    const uint32_t* f = flags[i];
    for (uint32_t spin = 0; spin < (1u << 23); ++spin) {
        if ((int32_t)(ld_acquire_sys(f) - target) >= 0) return;
        __nanosleep(500);
    }
    printf("orb_b200 xmap: timed out waiting for flag %d (target step %u, have %u)\n", i, target, ld_acquire_sys(f));
    asm volatile("trap;");
}
// lane i release-stores `value` into flags[i] (i < n); everything this stream did before is visible system-wide first
__global__ void xmap_signal_kernel(uint32_t* const* flags, int n, uint32_t value) {
    __threadfence_system();
    const int i = threadIdx.x;
    if (i < n) st_release_sys(flags[i], value);
}

struct XmapChunk { int a, b, tile0, tile1; };

// the part of the flattened (pair, query tile) list that `rank` handles; pairs in canonical order (a, then b != a)
static void xmap_plan(int n_maps, const int* rows, int world, int rank, std::vector<XmapChunk>& out) {
    out.clear();
    long long total = 0;
    for (int a = 0; a < n_maps; ++a) total += (long long)(n_maps - 1) * ceil_div(rows[a], kXmapTile);
    const long long lo = total * rank / world, hi = total * (rank + 1) / world;
    long long pos = 0;
    for (int a = 0; a < n_maps; ++a) {
        const int tiles = ceil_div(rows[a], kXmapTile);
        for (int b = 0; b < n_maps; ++b) {
            if (b == a) continue;
            const long long s = pos > lo ? pos : lo, e = pos + tiles < hi ? pos + tiles : hi;
            if (e > s && rows[b] >= 0) out.push_back({a, b, (int)(s - pos), (int)(e - pos)});
            pos += tiles;
        }
    }
}

}  // namespace orb

using namespace orb;

struct orbm_xmap {
    int device = 0, rank = 0, world = 1, n_maps = 0, rows_cap = 0, slots = 0;
    uint8_t* window = nullptr;
    size_t window_bytes = 0, off_packed = 0, off_results = 0, set_bytes = 0, res_rows = 0;
    uint8_t* peer[kXmapMaxWorld] = {};      // base of every rank's window in this process' address space
    bool ipc_opened[kXmapMaxWorld] = {};
    bool attached = false;
    uint32_t step = 0;
    uint8_t *scr_a = nullptr, *scr_b = nullptr;   // local expanded operands (rows_cap x 256 B each)
    int* scr_part = nullptr;                      // partial results of a split candidate scan (hamming_mma.cu)
    CUtensorMap map_a, map_b, map_bh;
    void** d_ptrs = nullptr;                      // device array of flag pointers for the wait / signal kernels
    void** h_ptrs = nullptr;                      // pinned staging for it
    int ptr_cursor = 0;
    static constexpr int kPtrSlots = 4096;

    XmapCtrl* ctrl(int r) const { return (XmapCtrl*)peer[r]; }
    uint8_t* packed(int r, int slot, int half) const { return peer[r] + off_packed + ((size_t)slot * 2 + half) * set_bytes; }
    int* result(int r, int slot, int pair_slot, int which) const {
        return (int*)(peer[r] + off_results) + (((size_t)slot * (n_maps - 1) + pair_slot) * 3 + which) * res_rows;
    }
};

// flag pointer lists are written into a ring of pinned host memory and copied with the stream, so that calls stay asynchronous
static int push_flags(orbm_xmap* x, const std::vector<void*>& v, void*** d_out, cudaStream_t st) {
    ORB_REQUIRE((int)v.size() <= 32, "too many flags for one wait");
    if (x->ptr_cursor + (int)v.size() > orbm_xmap::kPtrSlots) x->ptr_cursor = 0;
    void** h = x->h_ptrs + x->ptr_cursor;
    void** d = x->d_ptrs + x->ptr_cursor;
    for (size_t i = 0; i < v.size(); ++i) h[i] = v[i];
    ORB_CUDA_TRY(cudaMemcpyAsync(d, h, v.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
    x->ptr_cursor += (int)v.size();
    *d_out = d;
    return ORB_OK;
}
static int enqueue_wait(orbm_xmap* x, const std::vector<void*>& flags, uint32_t target, cudaStream_t st) {
    if (flags.empty()) return ORB_OK;
    void** d = nullptr;
    const int rc = push_flags(x, flags, &d, st);
    if (rc != ORB_OK) return rc;
    xmap_wait_kernel<<<1, 32, 0, st>>>((const uint32_t* const*)d, (int)flags.size(), target);
    count_launch();
    return ORB_OK;
}
static int enqueue_signal(orbm_xmap* x, const std::vector<void*>& flags, uint32_t value, cudaStream_t st) {
    if (flags.empty()) return ORB_OK;
    void** d = nullptr;
    const int rc = push_flags(x, flags, &d, st);
    if (rc != ORB_OK) return rc;
    xmap_signal_kernel<<<1, 32, 0, st>>>((uint32_t* const*)d, (int)flags.size(), value);
    count_launch();
    return ORB_OK;
}

extern "C" {

int orbm_xmap_plan(int n_maps, const int32_t* rows_per_map, int world, int rank, int32_t* chunks, int cap, int* n_chunks) {
    ORB_REQUIRE(n_maps >= 1 && world >= 1 && rank >= 0 && rank < world && rows_per_map && n_chunks, "bad plan arguments");
    for (int m = 0; m < n_maps; ++m) ORB_REQUIRE(rows_per_map[m] >= 0, "negative row count");
    std::vector<XmapChunk> v;
    xmap_plan(n_maps, rows_per_map, world, rank, v);
    *n_chunks = (int)v.size();
    for (int i = 0; i < (int)v.size() && i < cap; ++i) {
        if (!chunks) break;
        chunks[4 * i] = v[i].a; chunks[4 * i + 1] = v[i].b; chunks[4 * i + 2] = v[i].tile0; chunks[4 * i + 3] = v[i].tile1;
    }
    return (int)v.size() > cap && chunks ? ORB_ECAPACITY : ORB_OK;
}

int orbm_xmap_create(int device, int rank, int world, int n_maps, int rows_cap, orbm_xmap_t* out) {
    ORB_REQUIRE(out, "null output");
    *out = nullptr;
    ORB_REQUIRE(world >= 1 && world <= kXmapMaxWorld && rank >= 0 && rank < world, "world must be 1..16 and 0 <= rank < world");
    ORB_REQUIRE(n_maps >= 2 && ceil_div(n_maps, world) <= kXmapMaxSlots, "n_maps must be >= 2 and at most 8 per rank");
    ORB_REQUIRE(rows_cap >= 1 && rows_cap <= (1 << 24), "rows_cap out of range");
    ORB_CUDA_TRY(cudaSetDevice(device));
    orbm_xmap* x = new orbm_xmap;
    x->device = device; x->rank = rank; x->world = world; x->n_maps = n_maps; x->rows_cap = rows_cap;
    x->slots = ceil_div(n_maps, world);
    x->res_rows = align_up((size_t)rows_cap, 128);
    x->set_bytes = x->res_rows * 32;
    x->off_packed = kXmapCtrlBytes;
    x->off_results = x->off_packed + (size_t)x->slots * 2 * x->set_bytes;
    x->window_bytes = x->off_results + (size_t)x->slots * (n_maps - 1) * 3 * x->res_rows * sizeof(int);
    auto fail = [&](const char* what, cudaError_t e) {
        set_error("%s failed: %s", what, cudaGetErrorString(e));
        cudaFree(x->window); cudaFree(x->scr_a); cudaFree(x->scr_b); cudaFree(x->scr_part); cudaFree(x->d_ptrs); cudaFreeHost(x->h_ptrs);
        delete x;
        return ORB_ECUDA;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&x->window, x->window_bytes)) != cudaSuccess) return fail("cudaMalloc(window)", e);
    if ((e = cudaMemset(x->window, 0, x->window_bytes)) != cudaSuccess) return fail("cudaMemset(window)", e);
    if ((e = cudaMalloc(&x->scr_a, x->res_rows * 256)) != cudaSuccess) return fail("cudaMalloc(operand A)", e);
    if ((e = cudaMalloc(&x->scr_b, x->res_rows * 256)) != cudaSuccess) return fail("cudaMalloc(operand B)", e);
    if ((e = cudaMalloc(&x->scr_part, mma_partial_ints(x->res_rows) * sizeof(int))) != cudaSuccess) return fail("cudaMalloc(partial results)", e);
    if ((e = cudaMemset(x->scr_a, 0, x->res_rows * 256)) != cudaSuccess) return fail("cudaMemset", e);
    if ((e = cudaMemset(x->scr_b, 0, x->res_rows * 256)) != cudaSuccess) return fail("cudaMemset", e);
    if ((e = cudaMalloc(&x->d_ptrs, orbm_xmap::kPtrSlots * sizeof(void*))) != cudaSuccess) return fail("cudaMalloc(flags)", e);
    if ((e = cudaMallocHost(&x->h_ptrs, orbm_xmap::kPtrSlots * sizeof(void*))) != cudaSuccess) return fail("cudaMallocHost(flags)", e);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail("cudaDeviceSynchronize", e);
    // every kernel of the step is loaded while the device is idle (see mma_preload_kernels)
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, xmap_wait_kernel)) != cudaSuccess) return fail("cudaFuncGetAttributes", e);
    if ((e = cudaFuncGetAttributes(&fa, xmap_signal_kernel)) != cudaSuccess) return fail("cudaFuncGetAttributes", e);
    int rc = mma_preload_kernels();
    if (rc != ORB_OK) { cudaFree(x->window); cudaFree(x->scr_a); cudaFree(x->scr_b); cudaFree(x->scr_part); cudaFree(x->d_ptrs); cudaFreeHost(x->h_ptrs); delete x; return rc; }
    rc = mma_encode_operand_map(&x->map_a, x->scr_a, (long long)x->res_rows, true);
    if (rc == ORB_OK) rc = mma_encode_operand_map(&x->map_b, x->scr_b, (long long)x->res_rows, false);
    if (rc == ORB_OK) rc = mma_encode_operand_map_half(&x->map_bh, x->scr_b, (long long)x->res_rows);
    if (rc != ORB_OK) { fail("tensor map", cudaSuccess); return rc; }
    x->peer[rank] = x->window;
    x->attached = world == 1;
    *out = x;
    return ORB_OK;
}

void orbm_xmap_destroy(orbm_xmap_t x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (x->ipc_opened[r]) cudaIpcCloseMemHandle(x->peer[r]);
    cudaFree(x->window); cudaFree(x->scr_a); cudaFree(x->scr_b); cudaFree(x->scr_part); cudaFree(x->d_ptrs); cudaFreeHost(x->h_ptrs);
    delete x;
}

int orbm_xmap_ipc_handle(orbm_xmap_t x, void* handle64) {
    ORB_REQUIRE(x && handle64, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == ORBM_XMAP_HANDLE_BYTES, "handle size");
    ORB_CUDA_TRY(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    ORB_CUDA_TRY(cudaIpcGetMemHandle(&h, x->window));
    memcpy(handle64, &h, sizeof(h));
    return ORB_OK;
}

int orbm_xmap_attach_ipc(orbm_xmap_t x, const void* handles) {
    ORB_REQUIRE(x && handles, "null pointer");
    ORB_REQUIRE(!x->attached || x->world == 1, "already attached");
    ORB_CUDA_TRY(cudaSetDevice(x->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * ORBM_XMAP_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        ORB_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer[r] = (uint8_t*)p;
        x->ipc_opened[r] = true;
    }
    x->attached = true;
    return ORB_OK;
}

int orbm_xmap_attach_local(orbm_xmap_t* ctxs, int world) {
    ORB_REQUIRE(ctxs && world >= 1 && world <= kXmapMaxWorld, "bad context list");
    for (int r = 0; r < world; ++r) {
        ORB_REQUIRE(ctxs[r] && ctxs[r]->rank == r && ctxs[r]->world == world, "contexts must be given in rank order");
        ORB_REQUIRE(ctxs[r]->n_maps == ctxs[0]->n_maps && ctxs[r]->rows_cap == ctxs[0]->rows_cap, "contexts differ in shape");
    }
    for (int r = 0; r < world; ++r) {
        ORB_CUDA_TRY(cudaSetDevice(ctxs[r]->device));
        for (int q = 0; q < world; ++q) {
            if (q == r) continue;
            if (ctxs[q]->device != ctxs[r]->device) {
                int can = 0;
                ORB_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device));
                ORB_REQUIRE(can, "GPUs of the contexts cannot access each other's memory");
                const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ORB_CUDA_TRY(e);
                cudaGetLastError();
            }
            ctxs[r]->peer[q] = ctxs[q]->window;
        }
        ctxs[r]->attached = true;
    }
    return ORB_OK;
}

int orbm_knn2_allgather(orbm_xmap_t x, const uint8_t* const* d_sets, const int32_t* rows_per_map, void* stream) {
    ORB_REQUIRE(x && rows_per_map, "null pointer");
    ORB_REQUIRE(x->attached, "attach the peers' windows first (orbm_xmap_attach_ipc / orbm_xmap_attach_local)");
    for (int m = 0; m < x->n_maps; ++m) ORB_REQUIRE(rows_per_map[m] >= 0 && rows_per_map[m] <= x->rows_cap, "map larger than rows_cap");
    cudaStream_t st = (cudaStream_t)stream;
    ORB_CUDA_TRY(cudaSetDevice(x->device));
    const uint32_t step = ++x->step;
    const int half = step & 1;
    int rc;
    // ---- publish the maps this rank owns (map m: rank m % world, slot m / world) --------------------------------------
    int owned = 0;
    for (int m = x->rank; m < x->n_maps; m += x->world, ++owned) {
        ORB_REQUIRE(d_sets && (rows_per_map[m] == 0 || d_sets[owned]), "missing descriptor set of an owned map");
    }
    if (owned) {
        if (step >= 3) {   // the half being overwritten was last read in step - 2
            std::vector<void*> f;
            for (int r = 0; r < x->world; ++r) f.push_back(&x->ctrl(x->rank)->done[r]);
            if ((rc = enqueue_wait(x, f, step - 2, st)) != ORB_OK) return rc;
        }
        std::vector<void*> ready;
        for (int s = 0; s < owned; ++s) {
            const int m = x->rank + s * x->world;
            if (rows_per_map[m] > 0)
                ORB_CUDA_TRY(cudaMemcpyAsync(x->packed(x->rank, s, half), d_sets[s], (size_t)rows_per_map[m] * 32, cudaMemcpyDeviceToDevice, st));
            ready.push_back(&x->ctrl(x->rank)->ready[s]);
        }
        if ((rc = enqueue_signal(x, ready, step, st)) != ORB_OK) return rc;
    }
    // ---- this rank's share of the (pair, query tile) list ---------------------------------------------------------------
    std::vector<XmapChunk> plan;
    xmap_plan(x->n_maps, rows_per_map, x->world, x->rank, plan);
    int cur_a = -1, cur_t0 = -1, cur_t1 = -1, cur_b = -1;
    for (const XmapChunk& c : plan) {
        const int oa = c.a % x->world, sa = c.a / x->world, ob = c.b % x->world, sb = c.b / x->world;
        const int row0 = c.tile0 * kXmapTile;
        const int na = (c.tile1 * kXmapTile < rows_per_map[c.a] ? c.tile1 * kXmapTile : rows_per_map[c.a]) - row0;
        const int nb = rows_per_map[c.b];
        if (na <= 0) continue;
        std::vector<void*> need;
        const bool new_a = c.a != cur_a || c.tile0 != cur_t0 || c.tile1 != cur_t1, new_b = c.b != cur_b;
        if (new_a) need.push_back(&x->ctrl(oa)->ready[sa]);
        if (new_b && nb > 0) need.push_back(&x->ctrl(ob)->ready[sb]);
        if ((rc = enqueue_wait(x, need, step, st)) != ORB_OK) return rc;
        // the all-gather, fused into the operand expansion: packed rows are read out of the owner's window (NVLink when the
        // owner is another GPU) and land expanded in local memory
        if (new_a) {
            if ((rc = mma_expand_rows(x->packed(oa, sa, half) + (size_t)row0 * 32, na, true, x->scr_a, st)) != ORB_OK) return rc;
            cur_a = c.a; cur_t0 = c.tile0; cur_t1 = c.tile1;
        }
        if (new_b) {
            if ((rc = mma_expand_rows(x->packed(ob, sb, half), nb, false, x->scr_b, st)) != ORB_OK) return rc;
            cur_b = c.b;
        }
        const int ps = c.b < c.a ? c.b : c.b - 1;
        if ((rc = mma_launch_preexpanded(x->map_a, x->map_b, x->map_bh, na, nb, x->result(oa, sa, ps, 0) + row0, x->result(oa, sa, ps, 1) + row0,
                                         x->result(oa, sa, ps, 2) + row0, x->scr_part, st)) != ORB_OK)
            return rc;
    }
    // ---- done with every window for this step; owners then wait for their contributors -------------------------------------
    std::vector<void*> done;
    for (int r = 0; r < x->world; ++r) done.push_back(&x->ctrl(r)->done[x->rank]);
    if ((rc = enqueue_signal(x, done, step, st)) != ORB_OK) return rc;
    if (owned) {
        std::vector<void*> f;
        for (int r = 0; r < x->world; ++r) f.push_back(&x->ctrl(x->rank)->done[r]);
        if ((rc = enqueue_wait(x, f, step, st)) != ORB_OK) return rc;
    }
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbm_xmap_result(orbm_xmap_t x, int map_a, int map_b, const int32_t** d_idx, const int32_t** d_best, const int32_t** d_second) {
    ORB_REQUIRE(x && map_a >= 0 && map_a < x->n_maps && map_b >= 0 && map_b < x->n_maps && map_a != map_b, "bad map pair");
    ORB_REQUIRE(map_a % x->world == x->rank, "results of a pair live on the rank that owns the query map");
    const int slot = map_a / x->world, ps = map_b < map_a ? map_b : map_b - 1;
    if (d_idx) *d_idx = x->result(x->rank, slot, ps, 0);
    if (d_best) *d_best = x->result(x->rank, slot, ps, 1);
    if (d_second) *d_second = x->result(x->rank, slot, ps, 2);
    return ORB_OK;
}

}  // extern "C"
