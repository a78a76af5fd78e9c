// K11 + K12: the reference Frame's 64 x 48 keypoint lookup grid on the device, and the windowed
// nearest-neighbour search that walks it (no host round trip for the candidate gate).
//
//   K11 frame_grid_build   Frame::AssignFeaturesToGrid + PosInGrid     /root/reference/src/Frame.cc:230-245, 382-392
//   K12 window_knn2        Frame::GetFeaturesInArea                    src/Frame.cc:327-380
//                          + the best/second loop that consumes it     e.g. src/ORBmatcher.cc:84-116, 1385-1426
// The candidate ORDER of GetFeaturesInArea (cell column ix, then cell row iy, then insertion order inside the
// cell) decides which of two equally distant candidates wins; K12 reproduces it by giving every candidate its
// position in that order and reducing (distance, position) lexicographically. Stateless form only: searches whose
// gate depends on earlier matches take the distances (orbm_list_distances) and replay the loop on the host.
#include <new>

#include <algorithm>
#include <cstring>

#include "common.cuh"

struct orbm_grid {
    int device = 0, cap = 0;
    float min_x = 0, min_y = 0, w_inv = 0, h_inv = 0;
    int* d_cell_start = nullptr;  // [64*48 + 1]
    int* d_items = nullptr;       // [cap] keypoint indices, cell-major (ix-major, iy), insertion order inside a cell
    int n = 0;                    // keypoints of the last build (host copy not needed by the kernels)
    const orbx_keypoint* d_kps = nullptr;
    const int* d_count = nullptr;
    int* d_qcounts = nullptr;     // per-query candidate counts of orbm_window_lists_device, kept between calls: a stream-ordered
    int qcounts_cap = 0;          // allocation per call is given back at the call's own synchronisation and costs ~1 ms to get again
};

namespace orb {

constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;

__device__ __forceinline__ int c_round(float v) { return (int)roundf(v); }  // C round(): half away from zero

// single block: counting sort of the keypoints by cell (cell id = ix * 48 + iy, the reference's iteration order)
__global__ void __launch_bounds__(1024)
frame_grid_build_kernel(const orbx_keypoint* __restrict__ kps, const int* __restrict__ count, int cap, float min_x, float min_y,
                        float w_inv, float h_inv, int* __restrict__ cell_start, int* __restrict__ items) {
    __shared__ int s_cnt[kGridCells];
    __shared__ int s_warp[32];
    const int n = min(*count, cap);
    for (int i = threadIdx.x; i < kGridCells; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    auto cell_of = [&](int i) {
        const int px = c_round(__fmul_rn(__fsub_rn(kps[i].x, min_x), w_inv));
        const int py = c_round(__fmul_rn(__fsub_rn(kps[i].y, min_y), h_inv));
        return (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) ? -1 : px * kGridRows + py;
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const int c = cell_of(i); if (c >= 0) atomicAdd(&s_cnt[c], 1); }
    __syncthreads();
    // exclusive scan of the 3072 counters: 3 per thread
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int v[3], sum = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) { v[k] = s_cnt[threadIdx.x * 3 + k]; sum += v[k]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        s_warp[lane] = winc - w;
    }
    __syncthreads();
    int run = s_warp[warp] + inc - sum;
#pragma unroll
    for (int k = 0; k < 3; ++k) { cell_start[threadIdx.x * 3 + k] = run; s_cnt[threadIdx.x * 3 + k] = run; run += v[k]; }
    if (threadIdx.x == blockDim.x - 1) cell_start[kGridCells] = run;
    __syncthreads();
    // stable fill: one thread per cell walks the keypoints in index order would be O(n) per cell; instead every
    // keypoint finds its rank among EARLIER keypoints of the same cell (n is ~1-2 k, cells hold a handful)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = cell_of(i);
        if (c < 0) continue;
        int rank = 0;
        for (int j = 0; j < i; ++j) rank += cell_of(j) == c;
        items[s_cnt[c] + rank] = i;
    }
}

// one warp per query. Query q: centre (x, y), radius r, level window [minLevel, maxLevel] (GetFeaturesInArea's
// arguments), descriptor row q of A. Output: best / second distance, index of the first minimum, and the
// octaves of best and second (SearchByProjection(F, MPs) compares them, src/ORBmatcher.cc:119-122).
__global__ void __launch_bounds__(256)
window_knn2_kernel(const uint4* __restrict__ A, int nq, const float* __restrict__ qx, const float* __restrict__ qy,
                   const float* __restrict__ qr, const int* __restrict__ qmin, const int* __restrict__ qmax,
                   const orbx_keypoint* __restrict__ kps, const uint4* __restrict__ B, const int* __restrict__ cell_start,
                   const int* __restrict__ items, float min_x, float min_y, float w_inv, float h_inv,
                   int* __restrict__ o_idx, int* __restrict__ o_b1, int* __restrict__ o_b2, int* __restrict__ o_l1, int* __restrict__ o_l2) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float x = qx[q], y = qy[q], r = qr[q];
    const int minLevel = qmin[q], maxLevel = qmax[q];
    const bool check = minLevel > 0 || maxLevel >= 0;
    // cell window, exactly as src/Frame.cc:332-346
    const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, min_x), r), w_inv)));
    const int cx1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, min_x), r), w_inv)));
    const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, min_y), r), h_inv)));
    const int cy1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, min_y), r), h_inv)));
    const uint4 a0 = __ldg(A + (size_t)q * 2), a1 = __ldg(A + (size_t)q * 2 + 1);
    // per lane: best (dist, pos) and second dist with the octave that came with them
    uint32_t best = 0xffffffffu, second = 0xffffffffu;  // dist << 20 | pos  (pos < 2^20), octave kept aside
    int best_i = -1, best_l = -1, second_l = -1;
    int pos_base = 0;
    if (cx0 < kGridCols && cx1 >= 0 && cy0 < kGridRows && cy1 >= 0)
        for (int ix = cx0; ix <= cx1; ++ix) {
            // cells (ix, cy0..cy1) are contiguous in the cell-major item array
            const int lo = cell_start[ix * kGridRows + cy0], hi = cell_start[ix * kGridRows + cy1 + 1];
            for (int k = lo + lane; k < hi; k += 32) {
                const int j = items[k];
                const orbx_keypoint kp = kps[j];
                if (check && (kp.octave < minLevel || (maxLevel >= 0 && kp.octave > maxLevel))) continue;
                if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
                const uint4 b0 = __ldg(B + (size_t)j * 2), b1 = __ldg(B + (size_t)j * 2 + 1);
                const uint32_t d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                   __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                const uint32_t key = d << 20 | (uint32_t)(pos_base + k - lo);
                if (key < best) { second = best; second_l = best_l; best = key; best_i = j; best_l = kp.octave; }
                else if (key < second) { second = key; second_l = kp.octave; }
            }
            pos_base += hi - lo;
        }
    // merge lanes: the two smallest keys overall (keys are unique: positions differ)
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const int obi = __shfl_xor_sync(0xffffffffu, best_i, o), obl = __shfl_xor_sync(0xffffffffu, best_l, o);
        const int osl = __shfl_xor_sync(0xffffffffu, second_l, o);
        if (ob < best) {
            // other's best wins; second = min(my best, other's second)
            if (best < os) { second = best; second_l = best_l; } else { second = os; second_l = osl; }
            best = ob; best_i = obi; best_l = obl;
        } else {
            if (ob < second) { second = ob; second_l = obl; }
        }
    }
    if (lane == 0) {
        // the reference starts both distances at 256 and only accepts d < best (strict): a candidate at distance
        // 256 never enters, and ties go to the earlier position - both are what the (dist, pos) minimum yields
        const int d1 = best == 0xffffffffu ? 256 : (int)(best >> 20), d2 = second == 0xffffffffu ? 256 : (int)(second >> 20);
        const bool has1 = d1 < 256, has2 = d2 < 256;
        o_idx[q] = has1 ? best_i : -1; o_b1[q] = has1 ? d1 : 256; o_b2[q] = has2 ? d2 : 256;
        o_l1[q] = has1 ? best_l : -1; o_l2[q] = has2 ? second_l : -1;
    }
}

// Candidate lists of the same windows for the STATEFUL searches: pass 1 counts the in-window candidates of every
// query, pass 2 (after an exclusive scan of the counts) writes their indices in GetFeaturesInArea order together
// with the Hamming distances. The caller replays the reference's loop over (offsets, cands, dist).
template <bool kFill>
__global__ void __launch_bounds__(256)
window_lists_kernel(const uint4* __restrict__ A, int nq, const float* __restrict__ qx, const float* __restrict__ qy,
                    const float* __restrict__ qr, const int* __restrict__ qmin, const int* __restrict__ qmax,
                    const orbx_keypoint* __restrict__ kps, const uint4* __restrict__ B, const int* __restrict__ cell_start,
                    const int* __restrict__ items, float min_x, float min_y, float w_inv, float h_inv, int* __restrict__ counts,
                    const int* __restrict__ offsets, int* __restrict__ cands, int16_t* __restrict__ dist) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float x = qx[q], y = qy[q], r = qr[q];
    const int minLevel = qmin[q], maxLevel = qmax[q];
    const bool check = minLevel > 0 || maxLevel >= 0;
    const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, min_x), r), w_inv)));
    const int cx1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, min_x), r), w_inv)));
    const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, min_y), r), h_inv)));
    const int cy1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, min_y), r), h_inv)));
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (kFill) { a0 = __ldg(A + (size_t)q * 2); a1 = __ldg(A + (size_t)q * 2 + 1); }
    int written = kFill ? offsets[q] : 0;
    if (cx0 < kGridCols && cx1 >= 0 && cy0 < kGridRows && cy1 >= 0)
        for (int ix = cx0; ix <= cx1; ++ix) {
            const int lo = cell_start[ix * kGridRows + cy0], hi = cell_start[ix * kGridRows + cy1 + 1];
            for (int k0 = lo; k0 < hi; k0 += 32) {  // warp-uniform trip count: ballots keep the order
                const int k = k0 + lane;
                bool in = false;
                int j = 0;
                if (k < hi) {
                    j = items[k];
                    const orbx_keypoint kp = kps[j];
                    in = !(check && (kp.octave < minLevel || (maxLevel >= 0 && kp.octave > maxLevel))) &&
                         fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r;
                }
                const uint32_t ball = __ballot_sync(0xffffffffu, in);
                if (kFill && in) {
                    const int at = written + __popc(ball & ((1u << lane) - 1));
                    const uint4 b0 = __ldg(B + (size_t)j * 2), b1 = __ldg(B + (size_t)j * 2 + 1);
                    cands[at] = j;
                    dist[at] = (int16_t)(__popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w));
                }
                written += __popc(ball);
            }
        }
    if (!kFill && lane == 0) counts[q] = written;
}

// single block exclusive scan of counts[0..n) -> offsets[0..n]
__global__ void __launch_bounds__(1024) window_scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ offsets) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? counts[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < n) offsets[i] = carry + s_warp[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = s_carry;
}

}  // namespace orb

using namespace orb;

extern "C" {

int orbm_grid_create(int device, int capacity, orbm_grid_handle* out) {
    ORB_REQUIRE(out && capacity > 0 && capacity < (1 << 20), "bad grid capacity");
    *out = nullptr;
    ORB_CUDA_TRY(cudaSetDevice(device));
    orbm_grid* g = new (std::nothrow) orbm_grid();
    ORB_REQUIRE(g, "out of host memory");
    g->device = device; g->cap = capacity;
    cudaError_t e = cudaMalloc(&g->d_cell_start, (kGridCells + 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_items, (size_t)capacity * sizeof(int));
    if (e != cudaSuccess) { set_error("grid allocation failed: %s", cudaGetErrorString(e)); orbm_grid_destroy(g); return ORB_ECUDA; }
    *out = g;
    return ORB_OK;
}

void orbm_grid_destroy(orbm_grid_handle g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->d_cell_start) cudaFree(g->d_cell_start);
    if (g->d_items) cudaFree(g->d_items);
    if (g->d_qcounts) cudaFree(g->d_qcounts);
    delete g;
}

int orbm_grid_build_device(orbm_grid_handle g, const orbx_keypoint* d_kps, const int32_t* d_count, float min_x, float min_y, float max_x,
                           float max_y, void* stream) {
    ORB_REQUIRE(g && d_kps && d_count && max_x > min_x && max_y > min_y, "bad grid arguments");
    ORB_CUDA_TRY(cudaSetDevice(g->device));
    g->min_x = min_x; g->min_y = min_y;
    g->w_inv = (float)kGridCols / (max_x - min_x);  // mfGridElementWidthInv (src/Frame.cc:101-102)
    g->h_inv = (float)kGridRows / (max_y - min_y);
    g->d_kps = d_kps; g->d_count = d_count;
    frame_grid_build_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_kps, d_count, g->cap, g->min_x, g->min_y, g->w_inv, g->h_inv, g->d_cell_start,
                                                                 g->d_items);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbm_window_knn2_device(orbm_grid_handle g, const uint8_t* d_desc_frame, const uint8_t* d_queries, int nq, const float* d_x,
                            const float* d_y, const float* d_r, const int32_t* d_min_level, const int32_t* d_max_level, int32_t* d_idx,
                            int32_t* d_best, int32_t* d_second, int32_t* d_best_level, int32_t* d_second_level, void* stream) {
    ORB_REQUIRE(g && nq >= 0, "bad arguments");
    if (nq == 0) return ORB_OK;
    ORB_REQUIRE(g->d_kps && d_desc_frame && d_queries && d_x && d_y && d_r && d_min_level && d_max_level && d_idx && d_best && d_second &&
                    d_best_level && d_second_level,
                "null pointer / grid not built");
    ORB_CUDA_TRY(cudaSetDevice(g->device));
    window_knn2_kernel<<<ceil_div(nq * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)d_queries, nq, d_x, d_y, d_r, d_min_level, d_max_level, g->d_kps, (const uint4*)d_desc_frame, g->d_cell_start, g->d_items,
        g->min_x, g->min_y, g->w_inv, g->h_inv, d_idx, d_best, d_second, d_best_level, d_second_level);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbm_window_lists_device(orbm_grid_handle g, const uint8_t* d_desc_frame, const uint8_t* d_queries, int nq, const float* d_x,
                             const float* d_y, const float* d_r, const int32_t* d_min_level, const int32_t* d_max_level,
                             int32_t* d_offsets, int32_t* d_cands, int16_t* d_dist, int cap, int32_t* total_out, void* stream) {
    ORB_REQUIRE(g && nq >= 0 && total_out, "bad arguments");
    *total_out = 0;
    if (nq == 0) return ORB_OK;
    ORB_REQUIRE(g->d_kps && d_desc_frame && d_queries && d_x && d_y && d_r && d_min_level && d_max_level && d_offsets, "null pointer / grid not built");
    ORB_CUDA_TRY(cudaSetDevice(g->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (nq > g->qcounts_cap) {
        if (g->d_qcounts) { ORB_CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(g->d_qcounts); g->d_qcounts = nullptr; g->qcounts_cap = 0; }
        ORB_CUDA_TRY(cudaMalloc(&g->d_qcounts, (size_t)(nq + nq / 2 + 256) * sizeof(int)));
        g->qcounts_cap = nq + nq / 2 + 256;
    }
    int* d_counts = g->d_qcounts;
    const int blocks = ceil_div(nq * 32, 256);
    window_lists_kernel<false><<<blocks, 256, 0, st>>>((const uint4*)d_queries, nq, d_x, d_y, d_r, d_min_level, d_max_level, g->d_kps,
                                                       (const uint4*)d_desc_frame, g->d_cell_start, g->d_items, g->min_x, g->min_y, g->w_inv,
                                                       g->h_inv, d_counts, nullptr, nullptr, nullptr);
    window_scan_kernel<<<1, 1024, 0, st>>>(d_counts, nq, d_offsets);
    count_launch(2);
    ORB_CUDA_TRY(cudaGetLastError());
    int total = 0;
    ORB_CUDA_TRY(cudaMemcpyAsync(&total, d_offsets + nq, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORB_CUDA_TRY(cudaStreamSynchronize(st));  // the caller needs the total to size / check its buffers
    *total_out = total;
    if (total > cap) { set_error("window lists hold %d candidates, buffer has room for %d", total, cap); return ORB_ECAPACITY; }
    if (total == 0) return ORB_OK;
    ORB_REQUIRE(d_cands && d_dist, "null output");
    window_lists_kernel<true><<<blocks, 256, 0, st>>>((const uint4*)d_queries, nq, d_x, d_y, d_r, d_min_level, d_max_level, g->d_kps,
                                                      (const uint4*)d_desc_frame, g->d_cell_start, g->d_items, g->min_x, g->min_y, g->w_inv,
                                                      g->h_inv, nullptr, d_offsets, d_cands, d_dist);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// ---- host-buffer form ----------------------------------------------------------------------------------------------
// The facade's windowed searches on host-resident frames (std::vector<cv::KeyPoint> + cv::Mat): keypoints, descriptors and
// queries go up in ONE staged copy, grid build + window lists + distances run on the device, offsets / candidates /
// distances come back through pinned staging. Replaces a host loop of Frame::GetFeaturesInArea calls (src/Frame.cc:327-380;
// ~3 us each for the 100 px windows of SearchForInitialization) + DescriptorDistance per candidate. Per calling thread:
// stream, grid, device and pinned arenas (grown on demand, freed at thread exit).
extern "C++" {
namespace {
struct HostWindowContext {
    int device = -1;
    cudaStream_t st = nullptr;
    orbm_grid_handle grid = nullptr;
    int grid_cap = 0;
    uint8_t* dev = nullptr; size_t dev_cap = 0;
    uint8_t* pin = nullptr; size_t pin_cap = 0;
    ~HostWindowContext() { release(); }
    void release() {
        if (device >= 0) {
            cudaSetDevice(device);
            if (st) cudaStreamSynchronize(st);
            if (grid) orbm_grid_destroy(grid);
            if (dev) cudaFree(dev);
            if (pin) cudaFreeHost(pin);
            if (st) cudaStreamDestroy(st);
        }
        grid = nullptr; dev = nullptr; pin = nullptr; st = nullptr; dev_cap = pin_cap = 0; grid_cap = 0; device = -1;
    }
    int prepare(int d, int n_kps, size_t dev_bytes, size_t pin_bytes) {
        if (d != device) {
            release();
            ORB_CUDA_TRY(cudaSetDevice(d));
            ORB_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            device = d;
        } else {
            ORB_CUDA_TRY(cudaSetDevice(d));
        }
        if (n_kps > grid_cap) {
            if (grid) { cudaStreamSynchronize(st); orbm_grid_destroy(grid); grid = nullptr; grid_cap = 0; }
            const int want = std::max(4096, n_kps + n_kps / 2);
            int rc = orbm_grid_create(d, std::min(want, (1 << 20) - 1), &grid);
            if (rc) return rc;
            grid_cap = std::min(want, (1 << 20) - 1);
        }
        if (dev_bytes > dev_cap) {
            if (dev) { cudaStreamSynchronize(st); cudaFree(dev); dev = nullptr; dev_cap = 0; }
            ORB_CUDA_TRY(cudaMalloc(&dev, dev_bytes + dev_bytes / 2));
            dev_cap = dev_bytes + dev_bytes / 2;
        }
        if (pin_bytes > pin_cap) {
            if (pin) { cudaStreamSynchronize(st); cudaFreeHost(pin); pin = nullptr; pin_cap = 0; }
            ORB_CUDA_TRY(cudaMallocHost(&pin, pin_bytes + pin_bytes / 2));
            pin_cap = pin_bytes + pin_bytes / 2;
        }
        return ORB_OK;
    }
};
thread_local HostWindowContext tls_window;
}  // namespace
}  // extern "C++"

int orbm_window_lists(int device, const orbx_keypoint* kps, int n_kps, const float* bounds4, const uint8_t* desc_frame, const uint8_t* queries,
                      int nq, const float* x, const float* y, const float* r, const int32_t* min_level, const int32_t* max_level,
                      int32_t* offsets, int32_t* cands, int16_t* dist, int cap, int32_t* total_out) {
    ORB_REQUIRE(total_out && nq >= 0 && n_kps >= 0 && cap >= 0, "bad arguments");
    *total_out = 0;
    if (nq == 0) return ORB_OK;
    ORB_REQUIRE(offsets && bounds4 && queries && x && y && r && min_level && max_level && (n_kps == 0 || (kps && desc_frame)), "null pointer");
    ORB_REQUIRE(n_kps < (1 << 20), "too many keypoints");
    if (n_kps == 0) { for (int i = 0; i <= nq; ++i) offsets[i] = 0; return ORB_OK; }
    ORB_REQUIRE(cap == 0 || (cands && dist), "null output");
    auto pad = [](size_t b) { return align_up(b, 256); };
    // input block (one copy): count | kps | frame descriptors | query descriptors | x | y | r | min level | max level
    const size_t o_cnt = 0, o_kps = pad(4), o_df = o_kps + pad((size_t)n_kps * sizeof(orbx_keypoint)), o_dq = o_df + pad((size_t)n_kps * 32),
                 o_x = o_dq + pad((size_t)nq * 32), o_y = o_x + pad((size_t)nq * 4), o_r = o_y + pad((size_t)nq * 4), o_lo = o_r + pad((size_t)nq * 4),
                 o_hi = o_lo + pad((size_t)nq * 4), in_bytes = o_hi + pad((size_t)nq * 4);
    // output block: offsets | candidates | distances
    const size_t o_off = in_bytes, o_cd = o_off + pad((size_t)(nq + 1) * 4), o_ds = o_cd + pad((size_t)cap * 4), dev_bytes = o_ds + pad((size_t)cap * 2);
    HostWindowContext& c = tls_window;
    int rc;
    if ((rc = c.prepare(device, n_kps, dev_bytes, std::max(in_bytes, dev_bytes - in_bytes)))) return rc;
    uint8_t* p = c.pin;
    *reinterpret_cast<int32_t*>(p + o_cnt) = n_kps;
    memcpy(p + o_kps, kps, (size_t)n_kps * sizeof(orbx_keypoint));
    memcpy(p + o_df, desc_frame, (size_t)n_kps * 32);
    memcpy(p + o_dq, queries, (size_t)nq * 32);
    memcpy(p + o_x, x, (size_t)nq * 4); memcpy(p + o_y, y, (size_t)nq * 4); memcpy(p + o_r, r, (size_t)nq * 4);
    memcpy(p + o_lo, min_level, (size_t)nq * 4); memcpy(p + o_hi, max_level, (size_t)nq * 4);
    ORB_CUDA_TRY(cudaMemcpyAsync(c.dev, p, in_bytes, cudaMemcpyHostToDevice, c.st));
    uint8_t* d = c.dev;
    if ((rc = orbm_grid_build_device(c.grid, (const orbx_keypoint*)(d + o_kps), (const int32_t*)(d + o_cnt), bounds4[0], bounds4[1], bounds4[2],
                                     bounds4[3], c.st)))
        return rc;
    int total = 0;
    rc = orbm_window_lists_device(c.grid, d + o_df, d + o_dq, nq, (const float*)(d + o_x), (const float*)(d + o_y), (const float*)(d + o_r),
                                  (const int32_t*)(d + o_lo), (const int32_t*)(d + o_hi), (int32_t*)(d + o_off), (int32_t*)(d + o_cd),
                                  (int16_t*)(d + o_ds), cap, &total, c.st);
    *total_out = total;
    if (rc) return rc;   // ORB_ECAPACITY: *total_out tells how much room the lists need
    // results: offsets | used part of the candidates | used part of the distances -> pinned staging, one synchronisation
    uint8_t* q = c.pin;
    const size_t s_off = 0, s_cd = pad((size_t)(nq + 1) * 4), s_ds = s_cd + pad((size_t)total * 4);
    ORB_CUDA_TRY(cudaMemcpyAsync(q + s_off, d + o_off, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, c.st));
    if (total) {
        ORB_CUDA_TRY(cudaMemcpyAsync(q + s_cd, d + o_cd, (size_t)total * 4, cudaMemcpyDeviceToHost, c.st));
        ORB_CUDA_TRY(cudaMemcpyAsync(q + s_ds, d + o_ds, (size_t)total * 2, cudaMemcpyDeviceToHost, c.st));
    }
    ORB_CUDA_TRY(cudaStreamSynchronize(c.st));
    memcpy(offsets, q + s_off, (size_t)(nq + 1) * 4);
    if (total) { memcpy(cands, q + s_cd, (size_t)total * 4); memcpy(dist, q + s_ds, (size_t)total * 2); }
    return ORB_OK;
}

}  // extern "C"
