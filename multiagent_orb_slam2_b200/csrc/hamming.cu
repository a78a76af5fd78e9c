// Hamming matching kernels (K7 brute-force kNN-2, K8 candidate lists, distance matrix, ratio test).
//
// Replaces the DescriptorDistance loops of the reference matcher:
//   ORBmatcher::DescriptorDistance          /root/reference/src/ORBmatcher.cc:1649-1665
//   best/second selection rule              src/ORBmatcher.cc:104-116, 218-227, 449-458, 588-597
//   acceptance (threshold + ratio)          src/ORBmatcher.cc:229-232, 461-463, 600-603
//
// The selection rule "strict '<' updates in iteration order" is equivalent to
//   (best, second) = the two smallest values of the multiset {dist_j} U {256, 256},
//   idx            = first position attaining best,
// which is associative, so the database can be tiled / split freely and partial triples merged
// in position order (merge()). All arithmetic is integer: bit-exact by construction.
//
// Roofline: integer pipes. The plain form needs 8 x POPC.b32 per comparison (XU pipe, measured 15.3 results /
// clk / SM => 0.58 T cmp/s per B200); the carry-save form used here needs 4 POPC + 16 LOP3 and is bound by the
// 64-lane ALU pipe instead (0.71-0.87 T cmp/s measured). Memory traffic is (nA + nB) * 32 B per pass plus
// L2-resident tile re-reads, i.e. irrelevant next to the bit counting.
#include <atomic>

#include <cstring>

#include "common.cuh"

namespace orb {

struct Top2 {
    int b1, b2, pos;  // pos: iteration position of the first minimum (-1: none)
};

__device__ __forceinline__ void top2_update(Top2& t, int d, int pos) {
    t.b2 = min(t.b2, max(d, t.b1));
    t.pos = d < t.b1 ? pos : t.pos;
    t.b1 = min(t.b1, d);
}
__device__ __noinline__ void top2_update4(Top2& t, int d0, int d1, int d2, int d3, int pos) {
    top2_update(t, d0, pos); top2_update(t, d1, pos + 1); top2_update(t, d2, pos + 2); top2_update(t, d3, pos + 3);
}
// a precedes b in iteration order
__device__ __forceinline__ Top2 top2_merge(const Top2& a, const Top2& b) {
    Top2 r;
    r.b1 = min(a.b1, b.b1);
    r.pos = b.b1 < a.b1 ? b.pos : a.pos;
    r.b2 = min(max(a.b1, b.b1), min(a.b2, b.b2));
    return r;
}

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// 256-bit Hamming distance with a carry-save adder tree: 4 full adders (2 LOP3 each) fold the 8
// XOR words into {2 x ones, twos, fours}, so only 4 POPC (XU pipe, 16 lanes/clk/SM) are needed per
// comparison instead of 8; the extra LOP3s run on the 64-lane ALU pipe
// (tools/microbench/popc_bench.cu measures both forms).
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int hamming256_csa(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    const uint32_t s0 = lop3_xor3(x0, x1, x2), c0 = lop3_maj(x0, x1, x2);
    const uint32_t s1 = lop3_xor3(x3, x4, x5), c1 = lop3_maj(x3, x4, x5);
    const uint32_t s2 = lop3_xor3(s0, s1, x6), c2 = lop3_maj(s0, s1, x6);
    const uint32_t t0 = lop3_xor3(c0, c1, c2), f0 = lop3_maj(c0, c1, c2);
    return __popc(s2) + __popc(x7) + 2 * __popc(t0) + 4 * __popc(f0);
}

constexpr int kKnnThreads = 256;
constexpr int kTileRows = 256;  // database rows per shared-memory stage (8 KB)
constexpr int kStages = 2;

// position-explicit merge (partials may come in any order)
__device__ __forceinline__ Top2 top2_merge_any(const Top2& a, const Top2& b) {
    Top2 r;
    r.b1 = min(a.b1, b.b1);
    const bool take_b = b.b1 < a.b1 || (b.b1 == a.b1 && b.pos >= 0 && (a.pos < 0 || b.pos < a.pos));
    r.pos = take_b ? b.pos : a.pos;
    r.b2 = min(max(a.b1, b.b1), min(a.b2, b.b2));
    return r;
}

// grid: (query blocks, 1, pairs). A block of 256 threads serves 256/S queries; the S thread groups
// ("slices", warp-uniform) each scan 1/S of every database tile and are merged through shared
// memory, so that small problems still fill the machine (S=8: 32 queries per block).
// Pair p matches set pairs[2p] (queries) against set pairs[2p+1] (database); without a pair list
// both are set p. Set k lives at base + k*stride_rows*32 with count n_dev[k] (or n_const).
template <int S>
__global__ void __launch_bounds__(kKnnThreads)
knn2_kernel(const uint4* __restrict__ A, const int* __restrict__ nA_dev, int nA_const, int strideA_rows,
            const uint4* __restrict__ B, const int* __restrict__ nB_dev, int nB_const, int strideB_rows,
            const int* __restrict__ pairs, int out_stride,
            int* __restrict__ o_idx, int* __restrict__ o_b1, int* __restrict__ o_b2) {
    constexpr int QB = kKnnThreads / S;   // queries per block
    constexpr int R = kTileRows / S;      // rows of a tile per slice
    __shared__ __align__(128) uint4 tile[kStages][kTileRows * 2];
    __shared__ __align__(8) uint64_t full[kStages];
    __shared__ int m_b1[S > 1 ? S : 1][QB], m_b2[S > 1 ? S : 1][QB], m_pos[S > 1 ? S : 1][QB];

    const int pair = blockIdx.z;
    const int ia = pairs ? pairs[2 * pair] : pair, ib = pairs ? pairs[2 * pair + 1] : pair;
    const int nA = nA_dev ? min(nA_dev[ia], strideA_rows) : nA_const;  // device counts never exceed the set capacity
    const int nB = nB_dev ? min(nB_dev[ib], strideB_rows) : nB_const;
    const int q0 = blockIdx.x * QB;
    if (q0 >= nA) return;  // whole block exits together
    const int slice = threadIdx.x / QB, qi = threadIdx.x % QB;
    const int q = q0 + qi;
    const uint4* Ap = A + (size_t)ia * strideA_rows * 2;
    const uint4* Bp = B + (size_t)ib * strideB_rows * 2;
    const int ntiles = ceil_div(nB, kTileRows);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int r0 = t * kTileRows;
        const uint32_t bytes = (uint32_t)min(kTileRows, nB - r0) * 32u;
        mbar_expect_tx(&full[t % kStages], bytes);
        bulk_g2s(tile[t % kStages], Bp + (size_t)r0 * 2, bytes, &full[t % kStages]);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);

    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (q < nA) { a0 = __ldg(Ap + (size_t)q * 2); a1 = __ldg(Ap + (size_t)q * 2 + 1); }
    Top2 best{256, 256, -1};

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&full[s], (t / kStages) & 1);
        const int r0 = t * kTileRows;
        const int rows = min(kTileRows, nB - r0);
        const int jlo = slice * R, jhi = min(jlo + R, rows);
        const uint4* tp = tile[s];
        // four rows per trip; the (best, second) update runs only when one of the four can change it,
        // which is rare once the running second-best is small - a real branch, not predication
        int j = jlo;
        for (; j + 4 <= jhi; j += 4) {
            const int d0 = hamming256_csa(a0, a1, tp[2 * j], tp[2 * j + 1]);
            const int d1 = hamming256_csa(a0, a1, tp[2 * j + 2], tp[2 * j + 3]);
            const int d2 = hamming256_csa(a0, a1, tp[2 * j + 4], tp[2 * j + 5]);
            const int d3 = hamming256_csa(a0, a1, tp[2 * j + 6], tp[2 * j + 7]);
            if (min(min(d0, d1), min(d2, d3)) < best.b2) top2_update4(best, d0, d1, d2, d3, r0 + j);
        }
        for (; j < jhi; ++j) {
            const int d = hamming256_csa(a0, a1, tp[2 * j], tp[2 * j + 1]);
            if (d < best.b2) top2_update(best, d, r0 + j);
        }
        __syncthreads();  // every thread is done with stage s
        if (threadIdx.x == 0 && t + kStages < ntiles) issue(t + kStages);
    }
    if (S > 1) {
        m_b1[slice][qi] = best.b1; m_b2[slice][qi] = best.b2; m_pos[slice][qi] = best.pos;
        __syncthreads();
        if (slice == 0) {
#pragma unroll
            for (int k = 1; k < S; ++k) best = top2_merge_any(best, Top2{m_b1[k][qi], m_b2[k][qi], m_pos[k][qi]});
        }
    }
    if (slice == 0 && q < nA) {
        const size_t o = (size_t)pair * out_stride + q;
        o_idx[o] = best.pos; o_b1[o] = best.b1; o_b2[o] = best.b2;
    }
}

// One warp per query; lanes stride over the candidate list; positions (not row ids) break ties.
__global__ void knn2_lists_kernel(const uint4* __restrict__ A, int nA, const uint4* __restrict__ B,
                                  const int* __restrict__ offsets, const int* __restrict__ cands,
                                  int* __restrict__ o_idx, int* __restrict__ o_b1, int* __restrict__ o_b2) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nA) return;
    const uint4 a0 = __ldg(A + (size_t)q * 2), a1 = __ldg(A + (size_t)q * 2 + 1);
    const int lo = offsets[q], hi = offsets[q + 1];
    Top2 t{256, 256, -1};
    for (int k = lo + lane; k < hi; k += 32) {
        const int j = cands[k];
        top2_update(t, hamming256(a0, a1, __ldg(B + (size_t)j * 2), __ldg(B + (size_t)j * 2 + 1)), k);
    }
    // lanes hold interleaved subsequences: order by position explicitly
    for (int off = 16; off; off >>= 1) {
        Top2 o{__shfl_xor_sync(0xffffffffu, t.b1, off), __shfl_xor_sync(0xffffffffu, t.b2, off),
               __shfl_xor_sync(0xffffffffu, t.pos, off)};
        Top2 r;
        r.b1 = min(t.b1, o.b1);
        const bool take_o = o.b1 < t.b1 || (o.b1 == t.b1 && o.pos >= 0 && (t.pos < 0 || o.pos < t.pos));
        r.pos = take_o ? o.pos : t.pos;
        r.b2 = min(max(t.b1, o.b1), min(t.b2, o.b2));
        t = r;
    }
    if (lane == 0) { o_idx[q] = t.pos >= 0 ? cands[t.pos] : -1; o_b1[q] = t.b1; o_b2[q] = t.b2; }
}

// distance of every (query, candidate) pair of a CSR candidate list: out[k] for k in [offsets[q], offsets[q+1])
__global__ void list_distances_kernel(const uint4* __restrict__ A, int nA, const uint4* __restrict__ B,
                                      const int* __restrict__ offsets, const int* __restrict__ cands, int16_t* __restrict__ out) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nA) return;
    const uint4 a0 = __ldg(A + (size_t)q * 2), a1 = __ldg(A + (size_t)q * 2 + 1);
    for (int k = offsets[q] + lane; k < offsets[q + 1]; k += 32) {
        const int j = cands[k];
        out[k] = (int16_t)hamming256(a0, a1, __ldg(B + (size_t)j * 2), __ldg(B + (size_t)j * 2 + 1));
    }
}

__global__ void ratio_filter_kernel(const int* __restrict__ idx, const int* __restrict__ b1, const int* __restrict__ b2,
                                    int n, int th, int inclusive, float ratio, int* __restrict__ match) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d1 = b1[i], d2 = b2[i];
    const bool th_ok = inclusive ? d1 <= th : d1 < th;
    const bool ok = th_ok && idx[i] >= 0 && (float)d1 < __fmul_rn(ratio, (float)d2);
    match[i] = ok ? idx[i] : -1;
}

// 16x16 output tile per block; both operand tiles staged in shared memory
__global__ void distance_matrix_kernel(const uint4* __restrict__ A, int nA, const uint4* __restrict__ B, int nB,
                                       int16_t* __restrict__ out) {
    __shared__ uint4 sa[16][2], sb[16][2];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
    const int t = ty * 16 + tx;
    if (t < 32) { const int r = blockIdx.y * 16 + (t >> 1); sa[t >> 1][t & 1] = r < nA ? __ldg(A + (size_t)r * 2 + (t & 1)) : make_uint4(0, 0, 0, 0); }
    else if (t < 64) { const int u = t - 32, r = blockIdx.x * 16 + (u >> 1); sb[u >> 1][u & 1] = r < nB ? __ldg(B + (size_t)r * 2 + (u & 1)) : make_uint4(0, 0, 0, 0); }
    __syncthreads();
    if (i < nA && j < nB) out[(size_t)i * nB + j] = (int16_t)hamming256(sa[ty][0], sa[ty][1], sb[tx][0], sb[tx][1]);
}

// ---- host launcher ----------------------------------------------------------------------------
int launch_knn2_mma(const uint8_t* dA, const int* d_nA, int nA_max, int strideA_rows, const uint8_t* dB, const int* d_nB, int nB_max,
                    int strideB_rows, const int* d_pairs, int pairs, int out_stride, int* d_idx, int* d_b1, int* d_b2, cudaStream_t st,
                    int variant);
static thread_local int t_knn2_backend = 0;        // per calling thread: 0 = by problem size, 1 = POPC kernel, 2 / 3 = tensor-core kernel, one CTA / a CTA pair per query tile
constexpr long long kMmaMinWork = 1ll << 19;      // comparisons per call from which the tensor-core path is used (measured break-even: ~700 x 700)

static int launch_knn2(const uint8_t* dA, const int* d_nA, int nA_max, int strideA_rows, const uint8_t* dB,
                       const int* d_nB, int nB_max, int strideB_rows, const int* d_pairs, int pairs, int* d_idx, int* d_b1,
                       int* d_b2, cudaStream_t st) {
    if (pairs <= 0 || nA_max <= 0) return ORB_OK;
    // large problems go to the tensor cores (hamming_mma.cu): same results, several times the POPC pipe's throughput
    const long long work = (long long)pairs * nA_max * nB_max;
    const int backend = t_knn2_backend;
    if (nB_max > 0 && (backend >= 2 || (backend == 0 && work >= kMmaMinWork && nB_max >= 64)))
        return launch_knn2_mma(dA, d_nA, nA_max, strideA_rows, dB, d_nB, nB_max, strideB_rows, d_pairs, pairs, strideA_rows, d_idx, d_b1, d_b2, st,
                               backend == 0 ? 0 : backend - 1);
    ORB_REQUIRE(pairs <= 65535, "more than 65535 set pairs in one call");   // grid.z
    // pick the slice count so that the grid covers the machine about twice
    int S = 1;
    while (S < 8 && ceil_div(nA_max, kKnnThreads / S) * pairs < 2 * kNumSMs) S *= 2;
    const dim3 grid(ceil_div(nA_max, kKnnThreads / S), 1, pairs);
#define ORB_KNN2_LAUNCH(SS)                                                                                         \
    knn2_kernel<SS><<<grid, kKnnThreads, 0, st>>>((const uint4*)dA, d_nA, nA_max, strideA_rows, (const uint4*)dB, d_nB, \
                                                   nB_max, strideB_rows, d_pairs, strideA_rows, d_idx, d_b1, d_b2)
    switch (S) {
        case 1: ORB_KNN2_LAUNCH(1); break;
        case 2: ORB_KNN2_LAUNCH(2); break;
        case 4: ORB_KNN2_LAUNCH(4); break;
        default: ORB_KNN2_LAUNCH(8); break;
    }
#undef ORB_KNN2_LAUNCH
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb

using namespace orb;

extern "C" {

int orbm_set_knn2_backend(int backend) {
    ORB_REQUIRE(backend >= 0 && backend <= 3, "backend must be 0 (auto), 1 (POPC), 2 (tensor cores, one CTA per query tile) or 3 (tensor cores, CTA pairs)");
    t_knn2_backend = backend;
    return ORB_OK;
}

int orbm_knn2_device(const uint8_t* dA, int nA, const uint8_t* dB, int nB, int32_t* d_idx, int32_t* d_best,
                     int32_t* d_second, void* stream) {
    ORB_REQUIRE(nA >= 0 && nB >= 0, "negative row count");
    ORB_REQUIRE(nA == 0 || (dA && d_idx && d_best && d_second), "null pointer");
    ORB_REQUIRE(nB == 0 || dB, "null database");
    ORB_REQUIRE(((uintptr_t)dA & 15) == 0 && ((uintptr_t)dB & 15) == 0, "descriptor arrays must be 16-byte aligned");
    return launch_knn2(dA, nullptr, nA, nA, dB, nullptr, nB, nB, nullptr, 1, d_idx, d_best, d_second, (cudaStream_t)stream);
}

int orbm_knn2_batched_device(const uint8_t* dA, const int32_t* d_nA, int strideA_rows, const uint8_t* dB,
                             const int32_t* d_nB, int strideB_rows, int pairs, int32_t* d_idx, int32_t* d_best,
                             int32_t* d_second, void* stream) {
    ORB_REQUIRE(pairs >= 0 && strideA_rows >= 0 && strideB_rows >= 0, "negative size");
    ORB_REQUIRE(dA && dB && d_idx && d_best && d_second, "null pointer");
    ORB_REQUIRE(((uintptr_t)dA & 15) == 0 && ((uintptr_t)dB & 15) == 0, "descriptor arrays must be 16-byte aligned");
    return launch_knn2(dA, d_nA, strideA_rows, strideA_rows, dB, d_nB, strideB_rows, strideB_rows, nullptr, pairs, d_idx, d_best,
                       d_second, (cudaStream_t)stream);
}

int orbm_knn2_pairs_device(const uint8_t* d_sets, const int32_t* d_counts, int stride_rows, const int32_t* d_pairs, int pairs,
                           int32_t* d_idx, int32_t* d_best, int32_t* d_second, void* stream) {
    ORB_REQUIRE(pairs >= 0 && stride_rows >= 0, "negative size");
    ORB_REQUIRE(d_sets && d_counts && d_pairs && d_idx && d_best && d_second, "null pointer");
    ORB_REQUIRE(((uintptr_t)d_sets & 15) == 0, "descriptor arrays must be 16-byte aligned");
    return launch_knn2(d_sets, d_counts, stride_rows, stride_rows, d_sets, d_counts, stride_rows, stride_rows, d_pairs, pairs, d_idx,
                       d_best, d_second, (cudaStream_t)stream);
}

int orbm_knn2_lists_device(const uint8_t* dA, int nA, const uint8_t* dB, const int32_t* d_offsets,
                           const int32_t* d_cands, int32_t* d_idx, int32_t* d_best, int32_t* d_second, void* stream) {
    ORB_REQUIRE(nA >= 0, "negative row count");
    if (nA == 0) return ORB_OK;
    ORB_REQUIRE(dA && dB && d_offsets && d_cands && d_idx && d_best && d_second, "null pointer");
    knn2_lists_kernel<<<ceil_div(nA * 32, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)dA, nA, (const uint4*)dB, d_offsets,
                                                                                d_cands, d_idx, d_best, d_second);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbm_list_distances_device(const uint8_t* dA, int nA, const uint8_t* dB, const int32_t* d_offsets, const int32_t* d_cands,
                               int16_t* d_out, void* stream) {
    ORB_REQUIRE(nA >= 0, "negative row count");
    if (nA == 0) return ORB_OK;
    ORB_REQUIRE(dA && dB && d_offsets && d_cands && d_out, "null pointer");
    list_distances_kernel<<<ceil_div(nA * 32, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)dA, nA, (const uint4*)dB, d_offsets, d_cands, d_out);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbm_ratio_filter_device(const int32_t* d_idx, const int32_t* d_best, const int32_t* d_second, int n, int th,
                             int inclusive, float ratio, int32_t* d_match, void* stream) {
    ORB_REQUIRE(n >= 0, "negative count");
    if (n == 0) return ORB_OK;
    ORB_REQUIRE(d_idx && d_best && d_second && d_match, "null pointer");
    ratio_filter_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_idx, d_best, d_second, n, th, inclusive, ratio, d_match);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbm_distance_matrix_device(const uint8_t* dA, int nA, const uint8_t* dB, int nB, int16_t* d_out, void* stream) {
    ORB_REQUIRE(nA >= 0 && nB >= 0, "negative row count");
    if (nA == 0 || nB == 0) return ORB_OK;
    ORB_REQUIRE(dA && dB && d_out, "null pointer");
    distance_matrix_kernel<<<dim3(ceil_div(nB, 16), ceil_div(nA, 16)), dim3(16, 16), 0, (cudaStream_t)stream>>>(
        (const uint4*)dA, nA, (const uint4*)dB, nB, d_out);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// ---- host-buffer wrappers: H2D, kernel, D2H, synchronise (per-thread arena + stream: common.cuh) ------------------
extern "C++" {
namespace {
// the three result arrays of a search lie back to back on the device: one copy into pinned staging, one synchronisation
static int download_three(HostCallWorkspace& ws, const int* d_o, size_t n, int32_t* idx, int32_t* best, int32_t* second) {
    uint8_t* p = nullptr;
    int rc;
    if ((rc = ws.pinned(n * 12, &p))) return rc;
    ORB_CUDA_TRY(cudaMemcpyAsync(p, d_o, n * 12, cudaMemcpyDeviceToHost, ws.st));
    ORB_CUDA_TRY(cudaStreamSynchronize(ws.st));
    memcpy(idx, p, n * 4);
    memcpy(best, p + n * 4, n * 4);
    memcpy(second, p + n * 8, n * 4);
    return ORB_OK;
}
}  // namespace
}  // extern "C++"

int orbm_knn2(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, int32_t* idx, int32_t* best, int32_t* second) {
    ORB_REQUIRE(nA >= 0 && nB >= 0, "negative row count");
    if (nA == 0) return ORB_OK;
    ORB_REQUIRE(A && idx && best && second && (nB == 0 || B), "null pointer");
    HostCallWorkspace& ws = host_call_workspace();
    int rc;
    if ((rc = ws.begin(device, ws.need((size_t)nA * 32) + ws.need((size_t)nB * 32) + ws.need((size_t)nA * 12)))) return rc;
    uint8_t* dA = ws.take<uint8_t>((size_t)nA * 32);
    uint8_t* dB = ws.take<uint8_t>((size_t)nB * 32);
    int* o = ws.take<int>((size_t)nA * 3);
    ORB_CUDA_TRY(cudaMemcpyAsync(dA, A, (size_t)nA * 32, cudaMemcpyHostToDevice, ws.st));
    if (nB) ORB_CUDA_TRY(cudaMemcpyAsync(dB, B, (size_t)nB * 32, cudaMemcpyHostToDevice, ws.st));
    if ((rc = orbm_knn2_device(dA, nA, dB, nB, o, o + nA, o + 2 * (size_t)nA, ws.st))) return rc;
    return download_three(ws, o, (size_t)nA, idx, best, second);
}

int orbm_knn2_lists(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, const int32_t* offsets,
                    const int32_t* cands, int32_t* idx, int32_t* best, int32_t* second) {
    ORB_REQUIRE(nA >= 0 && nB >= 0, "negative row count");
    if (nA == 0) return ORB_OK;
    ORB_REQUIRE(A && B && offsets && cands && idx && best && second, "null pointer");
    const int total = offsets[nA];
    ORB_REQUIRE(total >= 0, "bad offsets");
    HostCallWorkspace& ws = host_call_workspace();
    int rc;
    if ((rc = ws.begin(device, ws.need((size_t)nA * 32) + ws.need((size_t)nB * 32) + ws.need((size_t)(nA + 1) * 4) + ws.need((size_t)total * 4) +
                                   ws.need((size_t)nA * 12))))
        return rc;
    uint8_t* dA = ws.take<uint8_t>((size_t)nA * 32);
    uint8_t* dB = ws.take<uint8_t>((size_t)nB * 32);
    int* dOf = ws.take<int>((size_t)nA + 1);
    int* dC = ws.take<int>((size_t)total);
    int* o = ws.take<int>((size_t)nA * 3);
    ORB_CUDA_TRY(cudaMemcpyAsync(dA, A, (size_t)nA * 32, cudaMemcpyHostToDevice, ws.st));
    ORB_CUDA_TRY(cudaMemcpyAsync(dB, B, (size_t)nB * 32, cudaMemcpyHostToDevice, ws.st));
    ORB_CUDA_TRY(cudaMemcpyAsync(dOf, offsets, (size_t)(nA + 1) * 4, cudaMemcpyHostToDevice, ws.st));
    if (total) ORB_CUDA_TRY(cudaMemcpyAsync(dC, cands, (size_t)total * 4, cudaMemcpyHostToDevice, ws.st));
    if ((rc = orbm_knn2_lists_device(dA, nA, dB, dOf, dC, o, o + nA, o + 2 * (size_t)nA, ws.st))) return rc;
    return download_three(ws, o, (size_t)nA, idx, best, second);
}

int orbm_list_distances(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, const int32_t* offsets, const int32_t* cands,
                        int16_t* out) {
    ORB_REQUIRE(nA >= 0 && nB >= 0, "negative row count");
    if (nA == 0) return ORB_OK;
    ORB_REQUIRE(A && B && offsets && cands && out, "null pointer");
    const int total = offsets[nA];
    ORB_REQUIRE(total >= 0, "bad offsets");
    if (total == 0) return ORB_OK;
    HostCallWorkspace& ws = host_call_workspace();
    int rc;
    if ((rc = ws.begin(device, ws.need((size_t)nA * 32) + ws.need((size_t)nB * 32) + ws.need((size_t)(nA + 1) * 4) + ws.need((size_t)total * 4) +
                                   ws.need((size_t)total * 2))))
        return rc;
    uint8_t* dA = ws.take<uint8_t>((size_t)nA * 32);
    uint8_t* dB = ws.take<uint8_t>((size_t)nB * 32);
    int* dOf = ws.take<int>((size_t)nA + 1);
    int* dC = ws.take<int>((size_t)total);
    int16_t* dO = ws.take<int16_t>((size_t)total);
    ORB_CUDA_TRY(cudaMemcpyAsync(dA, A, (size_t)nA * 32, cudaMemcpyHostToDevice, ws.st));
    ORB_CUDA_TRY(cudaMemcpyAsync(dB, B, (size_t)nB * 32, cudaMemcpyHostToDevice, ws.st));
    ORB_CUDA_TRY(cudaMemcpyAsync(dOf, offsets, (size_t)(nA + 1) * 4, cudaMemcpyHostToDevice, ws.st));
    ORB_CUDA_TRY(cudaMemcpyAsync(dC, cands, (size_t)total * 4, cudaMemcpyHostToDevice, ws.st));
    if ((rc = orbm_list_distances_device(dA, nA, dB, dOf, dC, dO, ws.st))) return rc;
    ORB_CUDA_TRY(cudaMemcpyAsync(out, dO, (size_t)total * 2, cudaMemcpyDeviceToHost, ws.st));
    ORB_CUDA_TRY(cudaStreamSynchronize(ws.st));
    return ORB_OK;
}

int orbm_distance_matrix(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, int16_t* out) {
    ORB_REQUIRE(nA >= 0 && nB >= 0, "negative row count");
    if (nA == 0 || nB == 0) return ORB_OK;
    ORB_REQUIRE(A && B && out, "null pointer");
    HostCallWorkspace& ws = host_call_workspace();
    int rc;
    if ((rc = ws.begin(device, ws.need((size_t)nA * 32) + ws.need((size_t)nB * 32) + ws.need((size_t)nA * nB * 2)))) return rc;
    uint8_t* dA = ws.take<uint8_t>((size_t)nA * 32);
    uint8_t* dB = ws.take<uint8_t>((size_t)nB * 32);
    int16_t* dO = ws.take<int16_t>((size_t)nA * nB);
    ORB_CUDA_TRY(cudaMemcpyAsync(dA, A, (size_t)nA * 32, cudaMemcpyHostToDevice, ws.st));
    ORB_CUDA_TRY(cudaMemcpyAsync(dB, B, (size_t)nB * 32, cudaMemcpyHostToDevice, ws.st));
    if ((rc = orbm_distance_matrix_device(dA, nA, dB, nB, dO, ws.st))) return rc;
    ORB_CUDA_TRY(cudaMemcpyAsync(out, dO, (size_t)nA * nB * 2, cudaMemcpyDeviceToHost, ws.st));
    ORB_CUDA_TRY(cudaStreamSynchronize(ws.st));
    return ORB_OK;
}

}  // extern "C"
