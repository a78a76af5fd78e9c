// K3: FAST-9/16 detection per 30-px cell with per-cell threshold fallback and 3x3 non-maximum
// suppression. One thread block per (cell, frame).
//
// Replaces the cell loop of ComputeKeyPointsOctTree, /root/reference/src/ORBextractor.cc:789-829,
// i.e. cv::FAST(cell + 6 px halo, iniThFAST, true) with the minThFAST retry when the cell comes back
// empty (812-816). Semantics pinned in SURVEY.md Appendix A.3 / oracle/cvprim.h:
//   * score(p) = max over the 16 arcs of 9 ring pixels of min(|I(p) - I(q)|, one sign) - 1;
//     corner at threshold t  <=>  score >= t;
//   * scores exist only for the cell interior [3, tw-3) x [3, th-3); NMS compares against the
//     thresholded score map of THIS cell (0 outside), strict '>' on all 8 neighbours.
// Since the NMS test for a pixel with score s >= t reduces to "all neighbours < s", the local-maximum
// flag is threshold independent; the ini/min decision is a block-wide count.
//
// Work is compacted between phases so that the expensive paths run on dense thread sets instead of
// diverged warps (the first version spent ~310 instructions per pixel, 2/3 of them in diverged
// slow paths):
//   1. every interior pixel: 4-point rejection test (any arc of 9 contains ring pixel 0 or 8 and 4
//      or 12)                                     -> survivors appended to a shared-memory list
//   2. list entries: full 16-pixel arc test + exact score        -> score map, corner list
//   3. corner list: 3x3 maximum test                              -> per-row bit masks
//   4. masks -> row-major ordered slot list (reference order inside the cell).
#include "extract_kernels.cuh"

namespace orb {

constexpr int kFastThreads = 128;
constexpr int kFastWarps = kFastThreads / 32;

// ring pixel k at (dx,dy): OpenCV order, starting at (0,3) going through (3,0), (0,-3), (-3,0)
__device__ __forceinline__ int ring_offset(int k, int tp) {
    constexpr int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    return dy[k] * tp + dx[k];
}

// full arc test at threshold th: 0 = no corner, 1 = bright arc (centre brighter than 9 contiguous ring
// pixels by more than th), 2 = dark arc
__device__ __forceinline__ int fast_arc_test(const uint8_t* __restrict__ p, int tp, int th) {
    const int v = p[0];
    const int lo = v - th, hi = v + th;
    // sign bits shifted into the masks: one IADD + one funnel shift per comparison. The ring ends up
    // in reverse bit order, which does not matter for circular contiguity.
    uint32_t mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int r = p[ring_offset(k, tp)];
        mb = __funnelshift_l((uint32_t)(r - lo), mb, 1);  // r < lo
        md = __funnelshift_l((uint32_t)(hi - r), md, 1);  // r > hi
    }
    auto has_arc9 = [](uint32_t m) {
        m |= m << 16;
        uint32_t a = m & (m >> 1);
        a &= a >> 2;
        a &= a >> 4;
        a &= m >> 8;
        return a != 0;
    };
    return has_arc9(mb) ? 1 : (has_arc9(md) ? 2 : 0);
}

// exact score of a known corner: max over the 16 arcs of the minimum margin on the corner's side, - 1
// (the other side cannot hold a 9-arc at the same time, so it cannot win the maximum)
__device__ __forceinline__ int fast_score_side(const uint8_t* __restrict__ p, int tp, bool dark) {
    const int v = p[0];
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { const int r = p[ring_offset(k, tp)]; d[k] = dark ? r - v : v - r; }
    int m2[16], m4[16], best = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int m8 = min(m4[k], m4[(k + 4) & 15]);
        best = max(best, min(m8, d[(k + 8) & 15]));
    }
    return best - 1;
}

// Shared memory (dynamic): tile | score | masks | offsets | corner lists. The tile (fast_bw x max_th
// bytes: the largest cell of the handle + 15 bytes, because the innermost TMA coordinate must be a
// multiple of 16 bytes) arrives by one TMA box load per block; whatever lies beyond this cell's own
// tw x th window is simply not looked at.
__global__ void __launch_bounds__(kFastThreads)
fast_cells_kernel(const Geometry* __restrict__ g, const CellDesc* __restrict__ cells, const __grid_constant__ TmaMaps maps,
                  uint32_t* __restrict__ slots, int* __restrict__ cell_counts) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int max_th = g->max_th, tp = g->fast_bw;
    uint8_t* tile = smem;
    uint8_t* score = smem + (size_t)max_th * tp;
    uint32_t* mask_ini = reinterpret_cast<uint32_t*>(score + (size_t)max_th * tp);  // [max_th][2]
    uint32_t* mask_all = mask_ini + 2 * max_th;
    int* offs = reinterpret_cast<int*>(mask_all + 2 * max_th);                      // [2*max_th + 1]
    uint16_t* list1 = reinterpret_cast<uint16_t*>(offs + 2 * max_th + 2);           // survivors of the 4-point test
    uint16_t* list2 = list1 + (size_t)max_th * tp;                                  // corners
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_cnt[kFastWarps], s_n2, s_total_ini;

    const CellDesc c = cells[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelGeom& L = g->lv[c.level];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw = c.tw, th = c.th;
    const int dw = tw - 6, dh = th - 6;  // interior (detection) size
    int* count_out = cell_counts + (size_t)frame * g->ncells + blockIdx.x;
    if (dw <= 0 || dh <= 0) {
        if (threadIdx.x == 0) *count_out = 0;
        return;
    }
    if (threadIdx.x == 0) {
        s_n2 = 0; s_total_ini = 0;
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, (uint32_t)(max_th * tp));
        tma_load_3d(tile, &maps.m[c.level], c.x0 & ~15, c.y0, frame, &bar);  // innermost TMA coordinate: 16-byte granular
    }
    {   // meanwhile: clear the score map and the masks
        uint32_t* scorew = reinterpret_cast<uint32_t*>(score);
        for (int i = threadIdx.x; i < (max_th * tp) >> 2; i += kFastThreads) scorew[i] = 0u;
        for (int i = threadIdx.x; i < 4 * max_th; i += kFastThreads) mask_ini[i] = 0u;  // mask_ini + mask_all
    }
    __syncthreads();       // barrier init + cleared maps visible
    mbar_wait(&bar, 0);    // tile landed

    const int minTh = g->minTh, iniTh = g->iniTh;
    const int phase = c.x0 & 15;
    const uint8_t* t0 = tile + phase;  // pixel (x, y) of the cell tile at t0[y * tp + x]
    uint8_t* sc0 = score + phase;

    // ---- phase 1: 4-point rejection on every interior pixel ----------------------------------------
    // branch-free test, incremental addressing, survivors appended to a list private to the warp
    // (no atomics): warp w owns list1[w * seg .. (w+1) * seg)
    const int seg = (max_th * tp) / kFastWarps;
    uint16_t* mylist = list1 + warp * seg;
    int cnt = 0;
    for (int x0 = 3; x0 < tw - 3; x0 += 32) {  // one trip unless the cell is wider than 38 px
        const bool xin = x0 + lane < tw - 3;
        const int x = xin ? x0 + lane : tw - 4;  // idle lanes re-test the last column (no branch), masked below
        const uint8_t* p = t0 + (3 + warp) * tp + x;
        const int up = -3 * tp, dn = 3 * tp;
        int at = (3 + warp) * tp + x;
        for (int y = 3 + warp; y < th - 3; y += kFastWarps, p += kFastWarps * tp, at += kFastWarps * tp) {
            const int v = p[0], lo = v - minTh, hi = v + minTh;
            const int r0 = p[dn], r8 = p[up], r4 = p[3], r12 = p[-3];
            const bool keep = xin & ((max(min(r0, r8), min(r4, r12)) < lo) | (min(max(r0, r8), max(r4, r12)) > hi));
            const uint32_t ball = __ballot_sync(0xffffffffu, keep);
            if (keep) mylist[cnt + __popc(ball & ((1u << lane) - 1))] = (uint16_t)at;
            cnt += __popc(ball);
        }
    }
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();

    // ---- phase 2a: full 16-pixel arc test on the survivors -> corner list (bit 15 = dark side) -------
    int c0 = s_cnt[0], c1 = s_cnt[1], c2 = s_cnt[2], c3 = s_cnt[3];
    const int n1 = c0 + c1 + c2 + c3;
    for (int i = threadIdx.x; i < n1; i += kFastThreads) {
        int k = i, w = 0;
        if (k >= c0) { k -= c0; w = 1; if (k >= c1) { k -= c1; w = 2; if (k >= c2) { k -= c2; w = 3; } } }
        const int at = list1[w * seg + k];
        const int side = fast_arc_test(t0 + at, tp, minTh);
        if (side) list2[atomicAdd(&s_n2, 1)] = (uint16_t)(at | (side == 2 ? 0x8000 : 0));
    }
    __syncthreads();
    // ---- phase 2b: exact score of the corners, dense -----------------------------------------------
    const int n2 = s_n2;
    for (int i = threadIdx.x; i < n2; i += kFastThreads) {
        const int e = list2[i], at = e & 0x7fff;
        sc0[at] = (uint8_t)fast_score_side(t0 + at, tp, (e & 0x8000) != 0);
        list2[i] = (uint16_t)at;
    }
    __syncthreads();

    // ---- phase 3: 3x3 strict maximum on the corners -> bit masks per (row, 32-column chunk) --------
    int my_ini = 0;
    for (int i = threadIdx.x; i < n2; i += kFastThreads) {
        const int at = list2[i];
        const uint8_t* q = sc0 + at;
        const int s = q[0];
        const int m = max(max(max(q[-tp - 1], q[-tp]), max(q[-tp + 1], q[-1])), max(max(q[1], q[tp - 1]), max(q[tp], q[tp + 1])));
        if (m < s) {
            const int y = at / tp, x = at - y * tp;
            const int e = (y - 3) * 2 + ((x - 3) >> 5);
            const uint32_t bit = 1u << ((x - 3) & 31);
            atomicOr(&mask_all[e], bit);
            if (s >= iniTh) { atomicOr(&mask_ini[e], bit); ++my_ini; }
        }
    }
    if (my_ini) atomicAdd(&s_total_ini, my_ini);
    __syncthreads();
    const uint32_t* mask = s_total_ini > 0 ? mask_ini : mask_all;  // minThFAST retry of an empty cell

    // ---- phase 4: exclusive prefix over the entries in row-major order (warp 0), ordered write ------
    const int nent = dh * 2;
    if (warp == 0) {
        int carry = 0;
        for (int e0 = 0; e0 < nent; e0 += 32) {
            const int e = e0 + lane;
            const int cnt = e < nent ? __popc(mask[e]) : 0;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (e < nent) offs[e] = carry + inc - cnt;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) *count_out = min(carry, L.slot_cap);
    }
    __syncthreads();

    uint32_t* out = slots + (size_t)frame * g->slot_words + L.slot_off + (size_t)c.ordinal * L.slot_cap;
    for (int e = threadIdx.x; e < nent; e += kFastThreads) {
        uint32_t m = mask[e];
        int rank = offs[e];
        const int y = 3 + (e >> 1), xb = 3 + (e & 1) * 32;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int x = xb + b;
            if (rank < L.slot_cap)
                out[rank] = (uint32_t)(x + c.offx) | (uint32_t)(y + c.offy) << 12 | (uint32_t)sc0[y * tp + x] << 24;
            ++rank;
        }
    }
}

static size_t fast_smem_bytes(const Geometry& hg, int tp) {
    const size_t px = (size_t)hg.max_th * tp;
    return 2 * px + (size_t)hg.max_th * 4 * sizeof(uint32_t) + (2 * (size_t)hg.max_th + 2) * sizeof(int) + 2 * px * sizeof(uint16_t);
}

int launch_fast(const Geometry& hg, const DeviceBuffers& db, const TmaMaps& maps, int n, cudaStream_t st) {
    fast_cells_kernel<<<dim3(hg.ncells, n), kFastThreads, fast_smem_bytes(hg, hg.fast_bw), st>>>(db.geom, db.cells, maps, db.slots,
                                                                                               db.cell_counts);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
