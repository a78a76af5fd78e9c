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
// Keypoints leave the kernel in the reference's order (row-major inside the cell) through a
// ballot + prefix-sum compaction into the cell's slot list; cells are stitched in row-major order
// by the quadtree kernel.
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

// exact FAST score if the pixel is a corner at threshold th (th >= 1), else 0
__device__ __forceinline__ int fast_score_px(const uint8_t* __restrict__ p, int tp, int th) {
    const int v = p[0];
    const int lo = v - th, hi = v + th;
    // any arc of 9 contains ring pixel 0 or 8, and 4 or 12: cheap exact rejection
    const int r0 = p[ring_offset(0, tp)], r8 = p[ring_offset(8, tp)];
    const int r4 = p[ring_offset(4, tp)], r12 = p[ring_offset(12, tp)];
    const bool brightish = (r0 < lo || r8 < lo) && (r4 < lo || r12 < lo);
    const bool darkish = (r0 > hi || r8 > hi) && (r4 > hi || r12 > hi);
    if (!brightish && !darkish) return 0;
    int r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = p[ring_offset(k, tp)];
    uint32_t mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { mb |= (uint32_t)(r[k] < lo) << k; md |= (uint32_t)(r[k] > hi) << k; }
    auto has_arc9 = [](uint32_t m) {
        m |= m << 16;
        uint32_t a = m & (m >> 1);
        a &= a >> 2;
        a &= a >> 4;
        a &= m >> 8;
        return a != 0;
    };
    const bool cb = has_arc9(mb), cd = has_arc9(md);
    if (!cb && !cd) return 0;
    // signed margins on the winning side (both sides cannot hold a 9-arc at once)
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = cb ? v - r[k] : r[k] - v;
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
    return best - 1;  // >= th by construction
}

// dynamic shared memory: tile[max_th][tp] | score[max_th][tp] | masks
__global__ void __launch_bounds__(kFastThreads)
fast_cells_kernel(const Geometry* __restrict__ g, const CellDesc* __restrict__ cells, FrameSet fs,
                  const uint8_t* __restrict__ pyr, uint32_t* __restrict__ slots, int* __restrict__ cell_counts, int tp) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int max_th = g->max_th;
    uint8_t* tile = smem;
    uint8_t* score = smem + (size_t)max_th * tp;
    uint32_t* mask_ini = reinterpret_cast<uint32_t*>(score + (size_t)max_th * tp);  // [max_th][2]
    uint32_t* mask_all = mask_ini + 2 * max_th;
    int* offs = reinterpret_cast<int*>(mask_all + 2 * max_th);                      // [2*max_th + 1]
    __shared__ int s_total_ini;

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
    int spitch;
    const uint8_t* src = level_ptr(*g, fs, pyr, frame, c.level, &spitch);
    src += (size_t)c.y0 * spitch + c.x0;

    if (threadIdx.x == 0) s_total_ini = 0;
    for (int y = warp; y < th; y += kFastWarps)
        for (int x = lane; x < tp; x += 32) {
            tile[y * tp + x] = x < tw ? src[(size_t)y * spitch + x] : 0;
            score[y * tp + x] = 0;
        }
    __syncthreads();

    const int minTh = g->minTh, iniTh = g->iniTh;
    for (int y = 3 + warp; y < th - 3; y += kFastWarps)
        for (int x = 3 + lane; x < tw - 3; x += 32)
            score[y * tp + x] = (uint8_t)fast_score_px(tile + y * tp + x, tp, minTh);
    __syncthreads();

    // local maxima + per-(row, 32-column chunk) ballots
    const int nchunk = (dw + 31) >> 5;  // 1 or 2
    int my_ini = 0;
    for (int y = 3 + warp; y < th - 3; y += kFastWarps)
        for (int ch = 0; ch < nchunk; ++ch) {
            const int x = 3 + ch * 32 + lane;
            bool ismax = false;
            int s = 0;
            if (x < tw - 3) {
                const uint8_t* q = score + y * tp + x;
                s = q[0];
                if (s > 0) {
                    const int m = max(max(max(q[-tp - 1], q[-tp]), max(q[-tp + 1], q[-1])),
                                      max(max(q[1], q[tp - 1]), max(q[tp], q[tp + 1])));
                    ismax = m < s;
                }
            }
            const uint32_t ball = __ballot_sync(0xffffffffu, ismax);
            const uint32_t bini = __ballot_sync(0xffffffffu, ismax && s >= iniTh);
            if (lane == 0) {
                mask_all[(y - 3) * 2 + ch] = ball;
                mask_ini[(y - 3) * 2 + ch] = bini;
                my_ini += __popc(bini);
            }
        }
    if (lane == 0 && my_ini) atomicAdd(&s_total_ini, my_ini);
    __syncthreads();
    const uint32_t* mask = s_total_ini > 0 ? mask_ini : mask_all;  // minThFAST retry of an empty cell

    // exclusive prefix over (row, chunk) entries in row-major order: warp 0
    const int nent = dh * 2;
    if (warp == 0) {
        int carry = 0;
        for (int e0 = 0; e0 < nent; e0 += 32) {
            const int e = e0 + lane;
            const int cnt = (e < nent && (e & 1) < nchunk) ? __popc(mask[e]) : 0;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (e < nent) offs[e] = carry + inc - cnt;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) { offs[nent] = carry; *count_out = min(carry, L.slot_cap); }
    }
    __syncthreads();

    uint32_t* out = slots + (size_t)frame * g->slot_words + L.slot_off + (size_t)c.ordinal * L.slot_cap;
    for (int y = 3 + warp; y < th - 3; y += kFastWarps)
        for (int ch = 0; ch < nchunk; ++ch) {
            const int e = (y - 3) * 2 + ch;
            const uint32_t m = mask[e];
            if (m >> lane & 1) {
                const int x = 3 + ch * 32 + lane;
                const int rank = offs[e] + __popc(m & ((1u << lane) - 1));
                if (rank < L.slot_cap)
                    out[rank] = (uint32_t)(x + c.offx) | (uint32_t)(y + c.offy) << 12 | (uint32_t)score[y * tp + x] << 24;
            }
        }
}

int launch_fast(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int n, cudaStream_t st) {
    const int tp = (hg.max_tw + 3) & ~3;
    const size_t smem = 2 * (size_t)hg.max_th * tp + (size_t)hg.max_th * 4 * sizeof(uint32_t) + (2 * (size_t)hg.max_th + 1) * sizeof(int);
    fast_cells_kernel<<<dim3(hg.ncells, n), kFastThreads, smem, st>>>(db.geom, db.cells, fs, db.pyr, db.slots, db.cell_counts, tp);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
