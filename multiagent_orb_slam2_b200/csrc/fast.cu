// K3: FAST-9/16 detection per 30-px cell with per-cell threshold fallback and 3x3 non-maximum
// suppression. One WARP per cell; every warp walks a strided list of cells of one frame. A cell's
// tile (TMA box load) is dead once its scores are known, so the next cell's load is issued at that
// point and lands while the warp finishes phases 3-4: no block barrier anywhere, no wait for the
// tile after a warp's first cell, and 6.4 KB of shared memory per warp. A block is ONE warp (measured at B=512:
// 0.765 ms against 0.797 / 0.814 / 0.860 ms with 2 / 4 / 8 warps per block - a block's slots free up as soon
// as its own cells are done).
//
// Replaces the cell loop of ComputeKeyPointsOctTree, /root/reference/src/ORBextractor.cc:789-829,
// i.e. cv::FAST(cell + 6 px halo, iniThFAST, true) with the minThFAST retry when the cell comes back
// empty (812-816). Semantics pinned in SURVEY.md Appendix A.3 / oracle/cvprim.h:
//   * score(p) = max over the 16 arcs of 9 ring pixels of min(|I(p) - I(q)|, one sign) - 1;
//     corner at threshold t  <=>  score >= t;
//   * scores exist only for the cell interior [3, tw-3) x [3, th-3); NMS compares against the
//     thresholded score map of THIS cell (0 outside), strict '>' on all 8 neighbours.
// Since the NMS test for a pixel with score s >= t reduces to "all neighbours < s", the local-maximum
// flag is threshold independent; the ini/min decision is a warp-wide count.
//
// Phases (all warp-synchronous, work compacted between them so that the expensive part runs on dense
// lanes):
//   1. every interior pixel, 2 rows x 4 per lane from aligned 32-bit words of the tile: rejection test (any arc
//      of 9 contains ring pixel 0 or 8 and 4 or 12: one of each pair must differ from the centre by more than
//      the threshold) on the four bytes of a word at once (VABSDIFF4 + SWAR compare)
//      -> list of words with survivors (ballot + one store; all-flat warp steps skip half of the test) -> survivor list
//   2. survivors: the 16 ring pixels packed two per register (k, k+8), the 16 arc minima by two rounds
//      of VIMNMX3.S16x2; corner <=> score >= threshold   -> score map, corner list (in place)
//   3. corner list: 3x3 strict maximum                   -> per-row bit mask
//   (1-3 run at iniThFAST; a cell that comes back empty repeats them at minThFAST on the same tile)
//   4. prefix over the mask words, then the corner list again: rank = offset + bits below -> the
//      row-major ordered slot list (reference order inside the cell).
#include <algorithm>
#include <atomic>

#include "extract_kernels.cuh"

namespace orb {

// ring pixel k at (dx,dy): OpenCV order, starting at (0,3) going through (3,0), (0,-3), (-3,0)
__device__ __forceinline__ int ring_offset(int k, int tp) {
    constexpr int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    return dy[k] * tp + dx[k];
}

constexpr int kFastWarps = 1;
constexpr int kFastThreads = kFastWarps * 32;
constexpr int kFastMaxCellsPerWarp = 8;

// per-warp shared memory: tile | score (interior + 1 px ring only) | mask | offs | list
struct FastLayout { int tile_bytes, score_pitch, score_bytes, mask_words, mask_off, list_off, warp_bytes; };
__host__ __device__ inline FastLayout fast_layout(int max_tw, int max_th, int tp) {
    FastLayout f;
    f.tile_bytes = max_th * tp;                      // multiple of 16 (tp is)
    f.score_pitch = (int)align_up((size_t)max_tw - 4, 4);  // score of cell pixel (x, y) at [(y - 2) * pitch + x - 2]
    f.score_bytes = (int)align_up((size_t)f.score_pitch * (max_th - 4), 16);
    // the same region first holds phase 1a's word list: 4 bytes per (row, aligned 4-pixel group) item, rows in bands of 8
    const int wl_bytes = ((max_th - 6 + 7) >> 3) * 8 * (((max_tw - 6 + 3) >> 2) + 1) * 4;
    if (wl_bytes > f.score_bytes) f.score_bytes = (int)align_up((size_t)wl_bytes, 16);
    f.mask_words = (max_tw - 6 > 32 ? 2 : 1) * (max_th - 6);  // one word per interior row and 32 columns
    f.mask_off = f.tile_bytes + f.score_bytes;
    f.list_off = f.mask_off + (int)align_up((size_t)2 * f.mask_words * 4, 16);
    const int npx = (max_tw - 6) * (max_th - 6);
    f.warp_bytes = (int)align_up((size_t)f.list_off + 2 * (size_t)npx + 16, 128);
    return f;
}

__device__ __forceinline__ uint32_t swap16(uint32_t x) { return __byte_perm(x, x, 0x1032); }

// max over the 16 arcs of 9 contiguous ring pixels of the minimum margin on one side, from the ring
// packed two pixels per register. sgn = +1: ring above the centre (margin r - v), -1: below (v - r).
// Margins carry a bias of 256 so that both halves stay positive and one IMAD chain packs and
// subtracts: D[k] = (256 + sgn (r[k] - v)) | (256 + sgn (r[k+8] - v)) << 16.
__device__ __forceinline__ int fast_side_margin(const int (&r)[16], int v, int sgn) {
    const uint32_t c0 = (uint32_t)(256 - sgn * v) * 0x10001u;
    const int s16 = sgn * 65536;
    uint32_t D[10];
#pragma unroll
    for (int k = 0; k < 8; ++k) D[k] = (uint32_t)(r[k] * sgn + (r[k + 8] * s16 + (int)c0));
    D[8] = swap16(D[0]); D[9] = swap16(D[1]);
    uint32_t Q[14];
#pragma unroll
    for (int k = 0; k < 8; ++k) Q[k] = __vimin3_s16x2(D[k], D[k + 1], D[k + 2]);  // min of 3 contiguous, arcs k and k+8
#pragma unroll
    for (int k = 0; k < 6; ++k) Q[8 + k] = swap16(Q[k]);
    uint32_t R[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) R[k] = __vimin3_s16x2(Q[k], Q[k + 3], Q[k + 6]);  // min of 9 contiguous
    const uint32_t m = __vimax3_s16x2(__vimax3_s16x2(R[0], R[1], R[2]), __vimax3_s16x2(R[3], R[4], R[5]), __vmaxs2(R[6], R[7]));
    return max((int)(m & 0xffffu), (int)(m >> 16)) - 256;
}

// TP: the tile's row pitch in bytes when it is known at compile time (64 for every cell up to 49 px wide: all
// standard camera settings), 0 = read it from the geometry. A constant pitch turns the ring offsets and row
// strides into immediates.
template <int TP>
__global__ void __launch_bounds__(kFastThreads, 32 / kFastWarps)
fast_cells_kernel(const Geometry* __restrict__ g, const CellDesc* __restrict__ cells, const __grid_constant__ TmaMaps maps,
                  uint32_t* __restrict__ slots, int* __restrict__ cell_counts, int cell_first, int cell_end) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kFastWarps];
    pdl_release_dependents();   // the quadtree grid behind this launch may be scheduled as soon as every FAST block has started
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int max_th = g->max_th, tp = TP ? TP : g->fast_bw, tpw = tp >> 2;
    const FastLayout lay = fast_layout(g->max_tw, max_th, tp);
    const int T = lay.tile_bytes, sp = lay.score_pitch;
    uint8_t* wbase = smem + (size_t)warp * lay.warp_bytes;
    uint8_t* score = wbase + T;
    uint32_t* mask = reinterpret_cast<uint32_t*>(wbase + lay.mask_off);
    uint32_t* offs = mask + lay.mask_words;
    uint16_t* list = reinterpret_cast<uint16_t*>(wbase + lay.list_off);
    uint64_t* bar = &bars[warp];
    const int frame = blockIdx.y, ncells = g->ncells;
    const int stride = gridDim.x * kFastWarps;       // cells gw, gw + stride, ... belong to this warp
    const int gw = cell_first + blockIdx.x * kFastWarps + warp;   // this launch covers cells [cell_first, cell_end): all, or one level
    const int minTh = g->minTh, iniTh = g->iniTh;
    if (gw >= cell_end) return;

    // one elected lane per warp owns the mbarrier and issues the tile loads. The tile is dead once a
    // cell's scores are known, so the next cell's load is issued there and lands during phases 3-4.
    auto issue = [&](int ci) {
        if (ci >= cell_end) return;
        const CellDesc c = cells[ci];
        if (c.tw > 6 && c.th > 6) {
            mbar_expect_tx(bar, (uint32_t)T);
            tma_load_3d(wbase, &maps.m[c.level], c.x0 & ~15, c.y0, frame, bar);  // innermost TMA coordinate: 16-byte granular
        }
    };
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        pdl_wait();   // behind the last resize launch: the pyramid is complete from here on
        issue(gw);
    }
    uint32_t parity = 0;

    for (int ci = gw; ci < cell_end; ci += stride) {
        __syncwarp();  // every lane is done with the previous cell
        const CellDesc c = cells[ci];
        const LevelGeom& L = g->lv[c.level];
        const int tw = c.tw, th = c.th;
        const int dw = tw - 6, dh = th - 6;  // interior (detection) size
        const int wsh = dw > 32 ? 1 : 0;     // mask words per interior row: 1 or 2
        int* count_out = cell_counts + (size_t)frame * ncells + ci;
        if (dw <= 0 || dh <= 0) {
            if (lane == 0) { *count_out = 0; issue(ci + stride); }
            continue;
        }
        for (int i = lane; i < lay.mask_words; i += 32) mask[i] = 0u;   // (the score map is cleared after phase 1)
        __syncwarp();
        mbar_wait(bar, parity);
        parity ^= 1u;

        const uint8_t* tile = wbase;  // pixel (x, y) of the cell tile at tile[y * tp + phase + x]
        const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(tile);
        const int phase = c.x0 & 15;

        // ---- phase 1: 4-point rejection, 2 rows x 4 pixels (two aligned words, one above the other) per lane ----
        const int bs = phase + 3, be = bs + dw;  // interior byte columns [bs, be) of a tile row
        const int gfirst = bs >> 2, G = ((be - 1) >> 2) - gfirst + 1;
        // rows are taken in bands of 8: row pair rp = 4 * band + q holds rows 8 * band + q and 8 * band + q + 4, so that the
        // four pairs a warp step touches with one load instruction lie in consecutive rows (2-way instead of 4-way bank
        // conflicts at the tile's 16-word pitch)
        const int nitems = ((dh + 7) >> 3) * 4 * G;
        const uint32_t Ginv = c.ginv;  // ceil(65536 / G), from the host: floor(i / G) == i * Ginv >> 16 for i * G < 65536
        // The reference runs cv::FAST at iniThFAST and, only when the cell comes back empty, again at minThFAST
        // (809-816). Same here: at iniTh far fewer pixels survive the rejection test and reach phases 2-3, and
        // the local-maximum test does not depend on the threshold (header), so pass 0 yields exactly the
        // iniTh result; an empty cell repeats phases 1-3 at minTh on the same tile.
        int n2 = 0;
        for (int pass = 0;; ++pass) {
        const int thr = pass ? minTh : iniTh;
        // Rejection test on four pixels at once, bytes in place (no unpacking): a pixel survives when one of ring
        // pixels (0,8) AND one of (4,12) differ from it by more than thr, whatever the sign. That is slightly weaker
        // than the signed 4-point test (measured: 9.2 % instead of 7.6 % of the pixels survive at thr 20) but costs
        // 19 instead of 47 instructions per word: VABSDIFF4 and, per difference, |d| > thr as the top bit of
        // ((d & 0x7f) + 127 - thr) | d. Phase 2 decides exactly. (thr > 126: everything survives.)
        const uint32_t Kthr = (uint32_t)max(127 - thr, 0) * 0x01010101u, Kall = thr > 126 ? 0x80808080u : 0u;
        // vertical half (ring pixels 0 and 8) and horizontal half (4 and 12) of the test; pixel j of the word at bit 8 j + 7
        auto vert4 = [&](uint32_t C, uint32_t U, uint32_t Dn) {
            const uint32_t a = __vabsdiffu4(C, U), b = __vabsdiffu4(C, Dn);
            return ((((a & 0x7f7f7f7fu) + Kthr) | a | ((b & 0x7f7f7f7fu) + Kthr) | b) | Kall) & 0x80808080u;
        };
        auto horz4 = [&](uint32_t C, uint32_t Lw, uint32_t Rw) {
            const uint32_t W12 = __funnelshift_r(Lw, C, 8);   // ring pixel 12 (x - 3) of the four centres
            const uint32_t W4 = __funnelshift_r(C, Rw, 24);   // ring pixel 4  (x + 3)
            const uint32_t c = __vabsdiffu4(C, W12), d = __vabsdiffu4(C, W4);
            return ((((c & 0x7f7f7f7fu) + Kthr) | c | ((d & 0x7f7f7f7fu) + Kthr) | d) | Kall) & 0x80808080u;
        };
        // ---- phase 1a: words with at least one survivor -> word list (in the score region, dead until phase 2) ----
        // entry = hit bits (8 j + 7) | byte column of the word (bits 0-6) | row << 8 (bits 8-14): one ballot and one
        // predicated store per row instead of a scan and four stores; a warp step whose 256 pixels all fail the vertical
        // half (flat image regions) skips the horizontal half
        uint32_t* wl = reinterpret_cast<uint32_t*>(score);
        int nw = 0;
        for (int i0 = 0; i0 < nitems; i0 += 32) {
            const bool act = i0 + lane < nitems;
            const int i = act ? i0 + lane : nitems - 1;
            const int rp = (int)((uint32_t)i * Ginv >> 16), gg = i - rp * G;
            const int y = 3 + 8 * (rp >> 2) + (rp & 3), wcol = gfirst + gg;
            const uint32_t* wp = tile32 + y * tpw + wcol;
            // byte columns outside [bs, be) masked; either row may lie below the interior
            const int bcol0 = wcol << 2;
            const int lead = max(bs - bcol0, 0), trail = max(bcol0 + 4 - be, 0);
            const uint32_t vw = (0x80808080u << (8 * lead)) & (0x80808080u >> (8 * trail)) & (act ? 0xffffffffu : 0u);
            const uint32_t C0 = wp[0], C1 = wp[4 * tpw];
            uint32_t kw0 = vert4(C0, wp[-3 * tpw], wp[3 * tpw]) & (y < th - 3 ? vw : 0u);
            uint32_t kw1 = vert4(C1, wp[tpw], wp[7 * tpw]) & (y + 4 < th - 3 ? vw : 0u);
            if (!__any_sync(0xffffffffu, (kw0 | kw1) != 0u)) continue;
            kw0 &= horz4(C0, wp[-1], wp[1]);
            kw1 &= horz4(C1, wp[4 * tpw - 1], wp[4 * tpw + 1]);
            const uint32_t b0 = __ballot_sync(0xffffffffu, kw0 != 0u), b1 = __ballot_sync(0xffffffffu, kw1 != 0u);
            const uint32_t pos = (uint32_t)(bcol0 | y << 8);
            if (kw0) wl[nw + __popc(b0 & lt_mask)] = kw0 | pos;
            nw += __popc(b0);
            if (kw1) wl[nw + __popc(b1 & lt_mask)] = kw1 | (pos + (4u << 8));
            nw += __popc(b1);
        }
        __syncwarp();
        // ---- phase 1b: word list -> pixel list (byte column | row << 7), one scan per 32 surviving words ----
        int n1 = 0;
        for (int i0 = 0; i0 < nw; i0 += 32) {
            const uint32_t w = i0 + lane < nw ? wl[i0 + lane] : 0u;
            const int cnt = __popc(w & 0x80808080u);
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            uint16_t* dst = list + n1 + inc - cnt;
            const uint32_t e = (w & 127u) | ((w >> 8) & 127u) << 7;
            if (w & 0x00000080u) *dst++ = (uint16_t)e;
            if (w & 0x00008000u) *dst++ = (uint16_t)(e + 1);
            if (w & 0x00800000u) *dst++ = (uint16_t)(e + 2);
            if (w & 0x80000000u) *dst++ = (uint16_t)(e + 3);
            n1 += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncwarp();
        {   // the word list is dead: the region becomes the (zeroed) score map
            uint4* z = reinterpret_cast<uint4*>(score);
            for (int i = lane; i < lay.score_bytes >> 4; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();

        // ---- phase 2: exact score of the survivors on their possible side; corner <=> score >= thr ----
        n2 = 0;
        for (int i0 = 0; i0 < n1; i0 += 32) {
            const bool act = i0 + lane < n1;
            const int e = list[act ? i0 + lane : n1 - 1];
            const int at = (e >> 7 & 127) * tp + (e & 127);
            const uint8_t* p = tile + at;
            const int v = p[0];
            int r[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) r[q] = p[ring_offset(q, tp)];
            // which side(s) can hold an arc of 9: one of (0,8) and one of (4,12) beyond the threshold
            const bool fb = max(min(r[0], r[8]), min(r[4], r[12])) < v - thr;
            const bool fa = min(max(r[0], r[8]), max(r[4], r[12])) > v + thr;
            int s = fast_side_margin(r, v, fb ? -1 : 1) - 1;
            if (__any_sync(0xffffffffu, fb & fa)) {  // both sides passed the 4-point test: rare
                if (fb & fa) s = max(s, fast_side_margin(r, v, 1) - 1);
            }
            const bool corner = act & (s >= thr);
            __syncwarp();  // all entries of this step are read before the compacted list overwrites them
            const uint32_t ball = __ballot_sync(0xffffffffu, corner);
            if (corner) {
                score[((e >> 7) - 2) * sp + (e & 127) - phase - 2] = (uint8_t)s;
                list[n2 + __popc(ball & lt_mask)] = (uint16_t)e;
            }
            n2 += __popc(ball);
        }
        __syncwarp();

        // ---- phase 3: 3x3 strict maximum on the corners -> bit mask per (row, 32-column half) ----------
        int total = 0;
        for (int i0 = 0; i0 < n2; i0 += 32) {
            const bool act = i0 + lane < n2;
            const int e = list[act ? i0 + lane : n2 - 1];
            const int y = e >> 7, x = (e & 127) - phase;
            const uint8_t* q = score + (y - 2) * sp + x - 2;
            const int s = q[0];
            const int m = max(max(max(q[-sp - 1], q[-sp]), max(q[-sp + 1], q[-1])), max(max(q[1], q[sp - 1]), max(q[sp], q[sp + 1])));
            const bool keep = act & (m < s);
            if (keep) atomicOr(&mask[((y - 3) << wsh) + ((x - 3) >> 5)], 1u << ((x - 3) & 31));
            total += __popc(__ballot_sync(0xffffffffu, keep));
        }
        if (total > 0 || pass == 1 || iniTh <= minTh) break;  // minThFAST retry of an empty cell
        }
        __syncwarp();  // nobody reads the tile any more
        if (lane == 0) issue(ci + stride);

        // ---- phase 4: exclusive prefix over the mask words in row-major order, ordered write ----------
        const int nent = dh << wsh;
        int carry = 0;
        for (int e0 = 0; e0 < nent; e0 += 32) {
            const int w = e0 + lane;
            const int cnt = w < nent ? __popc(mask[w]) : 0;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (w < nent) offs[w] = carry + inc - cnt;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) *count_out = min(carry, L.slot_cap);
        __syncwarp();
        uint32_t* out = slots + (size_t)frame * g->slot_words + L.slot_off + (size_t)c.ordinal * L.slot_cap;
        for (int i = lane; i < n2; i += 32) {
            const int e = list[i];
            const int y = e >> 7, x = (e & 127) - phase;
            const int w = ((y - 3) << wsh) + ((x - 3) >> 5);
            const uint32_t bit = 1u << ((x - 3) & 31), mw = mask[w];
            if (mw & bit) {
                const int rank = offs[w] + __popc(mw & (bit - 1u));
                if (rank < L.slot_cap)
                    out[rank] = (uint32_t)(x + c.offx) | (uint32_t)(y + c.offy) << 12 | (uint32_t)score[(y - 2) * sp + x - 2] << 24;
            }
        }
    }
}

// level < 0: the cells of all levels in one launch; otherwise only that level's (small batches run the levels as parallel branches)
int launch_fast(const Geometry& hg, const DeviceBuffers& db, const TmaMaps& maps, int n, cudaStream_t st, int level, bool pdl) {
    // list entries hold the byte column in 7 bits and the row in 7 bits
    ORB_REQUIRE(hg.fast_bw <= 128 && hg.max_th <= 128 && hg.max_tw - 6 <= 64, "FAST cell larger than 64 x 122 pixels");
    const FastLayout lay = fast_layout(hg.max_tw, hg.max_th, hg.fast_bw);
    const size_t smem = (size_t)lay.warp_bytes * kFastWarps;
    static std::atomic<int> attr_bytes[64];  // per device: dynamic shared memory opted in so far (a racing second call only repeats the attribute set)
    int dev = 0;
    ORB_CUDA_TRY(cudaGetDevice(&dev));
    auto kernel = hg.fast_bw == 64 ? fast_cells_kernel<64> : fast_cells_kernel<0>;
    if (smem > 48 * 1024 && (dev >= 64 || attr_bytes[dev] < (int)smem)) {   // device indices beyond the cache: always opt in
        ORB_CUDA_TRY(cudaFuncSetAttribute(fast_cells_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ORB_CUDA_TRY(cudaFuncSetAttribute(fast_cells_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 64) attr_bytes[dev] = (int)smem;
    }
    // cells per warp: as many as keeps every SM's warp slots (about 30 one-warp blocks) busy, at most 8
    const int cell_first = level < 0 ? 0 : hg.lv[level].cell_begin;
    const int cell_count = level < 0 ? hg.ncells : hg.lv[level].cell_count;
    if (cell_count <= 0) return ORB_OK;
    const long long warps_wanted = (long long)kNumSMs * 32;
    int cpw = (int)std::min<long long>(kFastMaxCellsPerWarp, std::max<long long>(1, (long long)cell_count * n / warps_wanted));
    const int blocks_x = ceil_div(ceil_div(cell_count, cpw), kFastWarps);
    if (pdl)
        ORB_CUDA_TRY(launch_pdl(kernel, dim3(blocks_x, n), dim3(kFastThreads), smem, st, db.geom, db.cells, maps, db.slots, db.cell_counts, cell_first,
                                cell_first + cell_count));
    else
        kernel<<<dim3(blocks_x, n), kFastThreads, smem, st>>>(db.geom, db.cells, maps, db.slots, db.cell_counts, cell_first, cell_first + cell_count);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
