// K7t: Hamming kNN-2 on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// The Hamming distance of two 256-bit descriptors is a dot product in disguise: with the bits mapped to +1 / -1,
// dot(a, b) = 256 - 2 * hamming(a, b). The best / second-best loop of ORBmatcher (/root/reference/src/ORBmatcher.cc:566-603
// and every other `DescriptorDistance` loop, 1649-1665) over a whole descriptor set is then an int8 GEMM
// (M = queries, N = candidates, K = 256) followed by a row-wise top-2 - and sm_100a's int8 tensor throughput is an order of
// magnitude above the POPC pipe (0.58 T comparisons/s/GPU) even at 8 bytes per bit-octet.
//
//   expand_pairs_kernel / expand_rows_kernel : 32-byte descriptors -> 256 int8 (queries +-64, database +-1), row-major, K
//                       contiguous (the K-major operand layout)
//   knn2_mma_kernel   : one CTA = 128 query rows, two CTAs per SM; TMA (SWIZZLE_128B) streams 128-candidate tiles through a
//                       shared-memory ring, one elected thread issues 8 tcgen05.mma (128 x 128 x 32) per tile into a
//                       double-buffered accumulator in TMEM, 8 epilogue warps read it back (tcgen05.ld, 16-bit packed) and
//                       keep the two largest dot products (= two smallest distances, first minimum by candidate order) per
//                       row; in a long scan groups of 16 candidates are first tested against the running second best.
//   knn2_mma_pair_kernel : the same on CTA pairs (cta_group::2, M = 256): selectable, measured, not the default.
//   top2_merge_splits_kernel : folds the partial results when a long database is scanned by 2 - 4 CTAs per query tile
//                       (against wave quantisation); the result is the one-pass result.
#include <cuda.h>

#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace orb {

// ---- bits -> +-1 bytes ---------------------------------------------------------------------------------------------
// thread = one 32-bit word of a descriptor -> 32 output bytes. Bit b of the word lands in byte b; any fixed permutation
// of K is fine as long as both operands use the same one.
__global__ void __launch_bounds__(256) expand_pm1_kernel(const uint32_t* __restrict__ bits, long long n_words, uint4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint32_t w = bits[i];
    uint32_t o[8];
#pragma unroll
    for (int nib = 0; nib < 8; ++nib) {
        const uint32_t x = (w >> (4 * nib)) & 0xFu;
        const uint32_t t = (x * 0x00204081u) & 0x01010101u;  // bit k of the nibble -> bit 0 of byte k
        o[nib] = 0xFFFFFFFFu ^ (t * 0xFEu);                  // 1 -> 0x01 (+1), 0 -> 0xFF (-1)
    }
    out[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
    out[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ---- tcgen05 / TMA plumbing ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
// mbarrier wait that traps instead of hanging the GPU when a producer never arrives (a descriptor mistake must not cost
// a whole box): ~2^26 polls is seconds of wall clock, far beyond any legitimate wait here
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {  // one full warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {  // arrives on bar when all MMAs issued so far by this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, one thread issues for the CTA
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);  // start address, 16-byte units
    d |= (uint64_t)1u << 16;                      // leading byte offset (unused for swizzled K-major), 16-byte units
    d |= (uint64_t)(1024u >> 4) << 32;            // stride byte offset between 8-row groups
    d |= (uint64_t)1u << 46;                      // descriptor version of sm_100
    d |= (uint64_t)2u << 61;                      // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 32 lanes x 32 consecutive columns: thread t of the warp gets lane (base lane + t), registers = columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int kMmaM = 128, kMmaN = 256, kMmaKBytes = 256;
constexpr int kATileBytes = kMmaM * kMmaKBytes;      // 32 KB: two 128-byte K chunks of 128 rows
constexpr int kBTileBytes = kMmaN * kMmaKBytes;      // 64 KB: two 128-byte K chunks of 256 rows

// the 8 MMAs of one 128 x 256 x 256 tile: K chunk c (128 bytes, its own swizzled block), K step k (32 bytes inside it)
__device__ __forceinline__ void issue_tile_mmas(uint32_t a_smem, uint32_t b_smem, uint32_t tmem_d) {
    constexpr uint32_t idesc = idesc_i8(kMmaM, kMmaN);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = smem_desc_k_sw128(a_smem + c * (kMmaM * 128) + k * 32);
            const uint64_t db = smem_desc_k_sw128(b_smem + c * (kMmaN * 128) + k * 32);
            mma_i8(tmem_d, da, db, idesc, (c | k) ? 1u : 0u);
        }
    }
}

// ---- the pair-list expansion ---------------------------------------------------------------------------------------
// side 0 = query sets (A, expanded to +-64), side 1 = database sets (B, +-1); pair p uses set d_pairs[2p + side] (or p when
// d_pairs is null); output row p * stride_rows + r. Rows beyond the set's count are left as they are (never used as results).
// The factor 64 on the query side makes the accumulator 64 * dot = 128 * (128 - distance): its low 16 bits are the selection
// key's upper field already, so the epilogue needs one packed add per two candidates to form the keys.
__global__ void __launch_bounds__(256) expand_pairs_kernel(const uint32_t* __restrict__ dA, const int* __restrict__ d_nA, int nA_max, int strideA,
                                                           const uint32_t* __restrict__ dB, const int* __restrict__ d_nB, int nB_max, int strideB,
                                                           const int* __restrict__ d_pairs, uint4* __restrict__ outA, uint4* __restrict__ outB) {
    pdl_release_dependents();   // the matcher behind this launch sets itself up (TMEM, barriers) meanwhile and waits before its first load
    const int p = blockIdx.y, side = blockIdx.z;
    const int set = d_pairs ? d_pairs[2 * p + side] : p;
    const int n = side ? (d_nB ? min(d_nB[set], nB_max) : nB_max) : (d_nA ? min(d_nA[set], nA_max) : nA_max);
    const int stride = side ? strideB : strideA;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // word index inside the set
    if (i >= n * 8) return;
    const uint32_t w = (side ? dB : dA)[(size_t)set * stride * 8 + i];
    const uint32_t neg = side ? 0xFFFFFFFFu : 0xC0C0C0C0u, flip = side ? 0xFEu : 0x80u;  // -1 / +1 or -64 / +64
    uint32_t o[8];
#pragma unroll
    for (int nib = 0; nib < 8; ++nib) {
        const uint32_t x = (w >> (4 * nib)) & 0xFu;
        const uint32_t t = (x * 0x00204081u) & 0x01010101u;
        o[nib] = neg ^ (t * flip);
    }
    uint4* out = (side ? outB : outA) + ((size_t)p * stride * 8 + i) * 2;
    out[0] = make_uint4(o[0], o[1], o[2], o[3]);
    out[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ---- the matcher ---------------------------------------------------------------------------------------------------
constexpr int kTileN = 128;                   // candidates per tile: 128 x 128 x 256 per 8 MMAs
constexpr int kBStageBytes = kTileN * kMmaKBytes;  // 32 KB
constexpr int kMmaStages = 2;                 // B tiles in flight in shared memory
constexpr int kMmaThreads = 320;              // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue
constexpr int kMmaSmemBytes = kATileBytes + kMmaStages * kBStageBytes + 1024;  // 97 KB: two CTAs per SM
constexpr int kTmemCols = 2 * kTileN;         // double-buffered accumulator; two CTAs per SM share the 512 columns
constexpr int kPruneFromTile = 16;            // candidate tiles after which the epilogue first tests groups against the second best

__device__ __forceinline__ void issue_tile_mmas_n128(uint32_t a_smem, uint32_t b_smem, uint32_t tmem_d) {
    constexpr uint32_t idesc = idesc_i8(kMmaM, kTileN);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = smem_desc_k_sw128(a_smem + c * (kMmaM * 128) + k * 32);
            const uint64_t db = smem_desc_k_sw128(b_smem + c * (kTileN * 128) + k * 32);
            mma_i8(tmem_d, da, db, idesc, (c | k) ? 1u : 0u);
        }
    }
}

// 32 lanes x 64 consecutive columns, the low 16 bits of two adjacent columns packed per register
__device__ __forceinline__ void tmem_ld64_pack16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// Selection keys, 16 bit: (256 - distance) << 7 | (127 - column). The accumulator's low half is 64 * dot = 128 * (128 - distance);
// adding 16384 + 127 - column gives the key. Larger key = smaller distance, and among equal distances the smaller column.
// Two candidates per register (VIADD.16x2 / VIMNMX.U16x2), two registers of running (best, second) for ILP.
struct Top2Packed {
    uint32_t m1[2], m2[2];
    __device__ __forceinline__ void reset() { m1[0] = m1[1] = m2[0] = m2[1] = 0u; }
    // two registers (four candidates) at a time: 5 min/max instructions, the last one three-input (VIMNMX3.U16x2)
    __device__ __forceinline__ void push2(int slot, uint32_t ka, uint32_t kb) {
        const uint32_t hi = __vmaxu2(ka, kb), lo = __vminu2(ka, kb);
        const uint32_t t = __vminu2(m1[slot], hi);
        m1[slot] = __vmaxu2(m1[slot], hi);
        m2[slot] = __vimax3_u16x2(m2[slot], t, lo);
    }
    __device__ __forceinline__ void reduce(int& k1, int& k2) const {
        const int a1 = m1[0] & 0xFFFF, b1 = m1[0] >> 16, c1 = m1[1] & 0xFFFF, d1 = m1[1] >> 16;
        const int a2 = m2[0] & 0xFFFF, b2 = m2[0] >> 16, c2 = m2[1] & 0xFFFF, d2 = m2[1] >> 16;
        const int x1 = max(a1, b1), x2 = max(min(a1, b1), max(a2, b2));
        const int y1 = max(c1, d1), y2 = max(min(c1, d1), max(c2, d2));
        k1 = max(x1, y1);
        k2 = max(min(x1, y1), max(x2, y2));
    }
};

// One epilogue warp's update of its rows' running (best, second, index) with the 64 candidates (two per register, raw
// accumulator halves) of tile t that start at candidate col0; `valid` of them are real.
__device__ __forceinline__ void top2_tile_update(const uint32_t (&r)[32], int t, int col0, int valid, int& R1, int& R2, int& Ridx) {
    Top2Packed acc;
    acc.reset();
    auto bias = [](int i) { return (uint32_t)(16384 + 127 - 2 * i) | (uint32_t)(16384 + 127 - (2 * i + 1)) << 16; };
    if (valid >= 64 && t >= kPruneFromTile) {
        // Steady state of a long database scan: almost no candidate beats a row's running SECOND best any more, and
        // one that does not cannot change (best, second, index) - an equal distance leaves all three as they are.
        // So 16 candidates at a time are first only compared, as raw accumulator halves (64 * dot = 128 * (128 -
        // distance), signed 16 bit), against the row's second best: 4 three-input packed maxima and one vote per
        // group instead of 28 key-building and selection instructions; the exact path below runs only for a group
        // in which some row of the warp has a contender. Rows are independent, so skipping is exact.
        const uint32_t thr2 = (uint32_t)((R2 * 128 - 16384) & 0xffff) * 0x10001u;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
            const int b = 8 * gq;
            const uint32_t mx = __vimax3_s16x2(__vimax3_s16x2(r[b], r[b + 1], r[b + 2]), __vimax3_s16x2(r[b + 3], r[b + 4], r[b + 5]),
                                               __vimax3_s16x2(r[b + 6], r[b + 7], thr2));
            if (__any_sync(0xffffffffu, mx != thr2)) {
#pragma unroll
                for (int i = b; i < b + 8; i += 2) acc.push2((i >> 1) & 1, __vadd2(r[i], bias(i)), __vadd2(r[i + 1], bias(i + 1)));
            }
        }
    } else if (valid >= 64) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) acc.push2((i >> 1) & 1, __vadd2(r[i], bias(i)), __vadd2(r[i + 1], bias(i + 1)));
    } else {
        auto mask = [valid](int i) { return (2 * i < valid ? 0xFFFFu : 0u) | (2 * i + 1 < valid ? 0xFFFF0000u : 0u); };
#pragma unroll
        for (int i = 0; i < 32; i += 2)
            acc.push2((i >> 1) & 1, __vadd2(r[i], bias(i)) & mask(i), __vadd2(r[i + 1], bias(i + 1)) & mask(i + 1));
    }
    int k1, k2;
    acc.reduce(k1, k2);
    const int d1 = k1 >> 7, d2 = k2 >> 7;
    if (d1 > R1) {
        R2 = max(R1, d2);
        R1 = d1;
        Ridx = col0 + 127 - (k1 & 127);
    } else {
        R2 = max(R2, d1);
    }
}

// the two column halves of a row meet in shared memory (epilogue warps only: named barrier 1, 256 threads), half 0 writes
__device__ __forceinline__ void top2_merge_halves_and_store(int h, int row, bool row_valid, int R1, int R2, int Ridx, int* s_r1, int* s_r2, int* s_ri,
                                                            int* o_idx, int* o_b1, int* o_b2) {
    if (h == 1) { s_r1[row] = R1; s_r2[row] = R2; s_ri[row] = Ridx; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (h == 0 && row_valid) {
        const int B1 = s_r1[row], B2 = s_r2[row], Bi = s_ri[row];
        const int m1 = max(R1, B1);
        const int m2 = max(min(R1, B1), max(R2, B2));
        int idx;
        if (B1 > R1) idx = Bi;
        else if (B1 < R1) idx = Ridx;
        else idx = (Ridx < 0 || Bi < 0) ? max(Ridx, Bi) : min(Ridx, Bi);
        *o_idx = idx;
        *o_b1 = 256 - m1;
        *o_b2 = 256 - m2;
    }
}

__global__ void __launch_bounds__(kMmaThreads, 2)
knn2_mma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const int* __restrict__ d_nA, int nA_max,
                int strideA, const int* __restrict__ d_nB, int nB_max, int strideB, const int* __restrict__ d_pairs, int out_stride,
                int* __restrict__ out_idx, int* __restrict__ out_b1, int* __restrict__ out_b2, long long split_stride) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a, bar_full[kMmaStages], bar_empty[kMmaStages], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_r1[kMmaM], s_r2[kMmaM], s_ri[kMmaM];

    const int p = blockIdx.y, mtile = blockIdx.x;
    const int setA = d_pairs ? d_pairs[2 * p] : p, setB = d_pairs ? d_pairs[2 * p + 1] : p;
    const int na = d_nA ? min(d_nA[setA], nA_max) : nA_max;
    const int nb = d_nB ? min(d_nB[setB], nB_max) : nB_max;
    const int row0 = mtile * kMmaM;
    if (row0 >= na) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // gridDim.z > 1: this CTA scans only its share of the candidate tiles and writes a partial result (split_stride apart)
    int* o_idx = out_idx + (size_t)blockIdx.z * split_stride + (size_t)p * out_stride;
    int* o_b1 = out_b1 + (size_t)blockIdx.z * split_stride + (size_t)p * out_stride;
    int* o_b2 = out_b2 + (size_t)blockIdx.z * split_stride + (size_t)p * out_stride;
    if (nb <= 0) {  // no candidates: the initial state of the reference loop
        for (int r = threadIdx.x; r < kMmaM; r += kMmaThreads)
            if (row0 + r < na) { o_idx[row0 + r] = -1; o_b1[row0 + r] = 256; o_b2[row0 + r] = 256; }
        return;
    }
    const int all_tiles = (nb + kTileN - 1) / kTileN;
    const int tile0 = (int)((long long)blockIdx.z * all_tiles / gridDim.z);                       // first candidate tile of this CTA
    const int ntiles = (int)((long long)(blockIdx.z + 1) * all_tiles / gridDim.z) - tile0;       // and how many (t below is local)
    uint8_t* sa = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B blocks: 1024-byte aligned
    uint8_t* sb = sa + kATileBytes;

    if (warp == 1) tmem_alloc(&tmem_base_s, kTmemCols);
    if (threadIdx.x == 0) {
        mbar_init(&bar_a, 1);
        for (int s = 0; s < kMmaStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 8); }
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const int arow = p * strideA + row0, brow = p * strideB;
            pdl_wait();   // programmatic dependent of the expansion kernel: the operands are complete from here on
            mbar_expect_tx(&bar_a, kATileBytes);
            tma_load_2d(sa, &map_a, 0, arow, &bar_a);
            tma_load_2d(sa + kMmaM * 128, &map_a, 128, arow, &bar_a);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kMmaStages;
                if (t >= kMmaStages) mbar_wait_or_trap(&bar_empty[s], ((t / kMmaStages) & 1) ^ 1);
                uint8_t* dst = sb + s * kBStageBytes;
                mbar_expect_tx(&bar_full[s], kBStageBytes);
                tma_load_2d(dst, &map_b, 0, brow + (tile0 + t) * kTileN, &bar_full[s]);
                tma_load_2d(dst + kTileN * 128, &map_b, 128, brow + (tile0 + t) * kTileN, &bar_full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            mbar_wait_or_trap(&bar_a, 0);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kMmaStages, a = t & 1;
                if (t >= 2) mbar_wait_or_trap(&bar_tempty[a], ((t >> 1) & 1) ^ 1);  // the epilogue drained this accumulator
                mbar_wait_or_trap(&bar_full[s], (t / kMmaStages) & 1);
                tc_fence_after();
                issue_tile_mmas_n128(smem_u32(sa), smem_u32(sb + s * kBStageBytes), tmem_base + a * kTileN);
                tc_commit(&bar_empty[s]);   // the shared-memory slot is free once these MMAs have read it
                tc_commit(&bar_tfull[a]);   // and the accumulator is complete
            }
        }
    } else {
        // ===== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves of 64 =====
        const int q = warp & 3, h = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int R1 = 0, R2 = 0, Ridx = -1;  // running (256 - distance) of best / second: 0 = distance 256 = "none yet"
        for (int t = 0; t < ntiles; ++t) {
            const int a = t & 1;
            mbar_wait_or_trap(&bar_tfull[a], (t >> 1) & 1);
            tc_fence_after();
            const int col0 = (tile0 + t) * kTileN + h * 64;      // candidate index of this half's first column
            const int valid = nb - col0;               // columns of this half that are real candidates
            uint32_t r[32];
            tmem_ld64_pack16(tmem_base + ((uint32_t)(q * 32) << 16) + a * kTileN + h * 64, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[a]);  // values are in registers: the accumulator can be overwritten
            top2_tile_update(r, t, col0, valid, R1, R2, Ridx);
        }
        top2_merge_halves_and_store(h, row, row0 + row < na, R1, R2, Ridx, s_r1, s_r2, s_ri, o_idx + row0 + row, o_b1 + row0 + row, o_b2 + row0 + row);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- the matcher on CTA pairs (cta_group::2) ---------------------------------------------------------------------
// Two CTAs of a cluster (the two SMs of a TPC) run ONE tcgen05.mma of M = 256: each CTA holds its own 128 query rows and
// its own 128 x 128 accumulator in its TMEM, and stages only HALF of every candidate tile (64 rows) in its shared memory -
// the tensor cores exchange the halves. Per comparison that is 3/4 of the shared-memory operand reads and half of the TMA
// writes of the one-CTA kernel, whose steady state is bound by exactly that (182 B/clk wanted of 128: DESIGN.md section 4).
// Roles per CTA as above; only the leader (cluster rank 0) issues MMAs. Barriers: every TMA of either CTA reports to the
// LEADER's full barrier; the leader's commits are multicast to both CTAs' empty / accumulator-full barriers; both CTAs'
// epilogue warps arrive on the leader's accumulator-empty barrier.
constexpr int kPairHalfN = kTileN / 2;                                // candidates a CTA stages per tile
constexpr int kPairStageBytes = kPairHalfN * kMmaKBytes;              // 16 KB
constexpr int kPairStages = 4;                                        // 4 half tiles per CTA = 4 tiles in flight per pair
constexpr int kPairSmemBytes = kATileBytes + kPairStages * kPairStageBytes + 1024;  // 97 KB: two CTAs (of two pairs) per SM

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA box load into THIS CTA's shared memory that reports its bytes to the barrier at the same offset in the pair's even CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar) & 0xFEFFFFFFu)
                 : "memory");
}
// arrive on `bar` of CTA `cta` of the cluster. Default (CTA-scope) semantics on purpose: what the barrier orders here are TMEM
// reads, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync have already retired - the .release.cluster form costs a
// MEMBAR.ALL.GPU + ERRBAR per arrival (20 % of this kernel's stall samples when it was tried)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_or_trap(uint64_t* bar, uint32_t parity) {  // arrivals come from both CTAs
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t cols) {  // the same warp of both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {  // arrives on `bar` of BOTH CTAs when the MMAs issued so far are done
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMmaThreads, 2)
knn2_mma_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bh, const int* __restrict__ d_nA,
                     int nA_max, int strideA, const int* __restrict__ d_nB, int nB_max, int strideB, const int* __restrict__ d_pairs,
                     int out_stride, int* __restrict__ out_idx, int* __restrict__ out_b1, int* __restrict__ out_b2, long long split_stride) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a, bar_full[kPairStages], bar_empty[kPairStages], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_r1[kMmaM], s_r2[kMmaM], s_ri[kMmaM];

    const uint32_t rank = cluster_ctarank();     // == blockIdx.x & 1
    const bool leader = rank == 0;
    const int p = blockIdx.y;
    const int setA = d_pairs ? d_pairs[2 * p] : p, setB = d_pairs ? d_pairs[2 * p + 1] : p;
    const int na = d_nA ? min(d_nA[setA], nA_max) : nA_max;
    const int nb = d_nB ? min(d_nB[setB], nB_max) : nB_max;
    const int pair_row0 = (int)(blockIdx.x >> 1) * (2 * kMmaM);
    if (pair_row0 >= na) return;                 // both CTAs of the pair leave together
    const int row0 = pair_row0 + (int)rank * kMmaM;   // may lie beyond na for the odd CTA of the last pair: computes, writes nothing
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // gridDim.z > 1: this CTA scans only its share of the candidate tiles and writes a partial result (split_stride apart)
    int* o_idx = out_idx + (size_t)blockIdx.z * split_stride + (size_t)p * out_stride;
    int* o_b1 = out_b1 + (size_t)blockIdx.z * split_stride + (size_t)p * out_stride;
    int* o_b2 = out_b2 + (size_t)blockIdx.z * split_stride + (size_t)p * out_stride;
    if (nb <= 0) {
        for (int r = threadIdx.x; r < kMmaM; r += kMmaThreads)
            if (row0 + r < na) { o_idx[row0 + r] = -1; o_b1[row0 + r] = 256; o_b2[row0 + r] = 256; }
        return;
    }
    const int all_tiles = (nb + kTileN - 1) / kTileN;
    const int tile0 = (int)((long long)blockIdx.z * all_tiles / gridDim.z);                       // first candidate tile of this CTA
    const int ntiles = (int)((long long)(blockIdx.z + 1) * all_tiles / gridDim.z) - tile0;       // and how many (t below is local)
    uint8_t* sa = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // same offset in both CTAs
    uint8_t* sb = sa + kATileBytes;

    if (threadIdx.x == 0) {
        mbar_init(&bar_a, 1);
        for (int s = 0; s < kPairStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 16); }  // 8 epilogue warps of each CTA
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_pair(&tmem_base_s, kTmemCols);
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs initialised before either CTA's TMA or commit can reach them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own query rows, own half of every candidate tile =====
        if (lane == 0) {
            const int arow = p * strideA + row0, brow = p * strideB + (int)rank * kPairHalfN;
            pdl_wait();
            if (leader) mbar_expect_tx(&bar_a, 2 * kATileBytes);
            tma_load_2d_pair(sa, &map_a, 0, arow, &bar_a);
            tma_load_2d_pair(sa + kMmaM * 128, &map_a, 128, arow, &bar_a);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kPairStages;
                if (t >= kPairStages) mbar_wait_or_trap(&bar_empty[s], ((t / kPairStages) & 1) ^ 1);
                uint8_t* dst = sb + s * kPairStageBytes;
                if (leader) mbar_expect_tx(&bar_full[s], 2 * kPairStageBytes);
                tma_load_2d_pair(dst, &map_bh, 0, brow + (tile0 + t) * kTileN, &bar_full[s]);
                tma_load_2d_pair(dst + kPairHalfN * 128, &map_bh, 128, brow + (tile0 + t) * kTileN, &bar_full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA for the pair =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = idesc_i8(2 * kMmaM, kTileN);
            mbar_wait_or_trap(&bar_a, 0);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kPairStages, a = t & 1;
                if (t >= 2) mbar_wait_cluster_or_trap(&bar_tempty[a], ((t >> 1) & 1) ^ 1);  // both CTAs' epilogues drained accumulator a
                mbar_wait_or_trap(&bar_full[s], (t / kPairStages) & 1);                     // both halves of the tile have landed
                tc_fence_after();
                const uint32_t a_smem = smem_u32(sa), b_smem = smem_u32(sb + s * kPairStageBytes), tmem_d = tmem_base + a * kTileN;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_i8_pair(tmem_d, smem_desc_k_sw128(a_smem + c * (kMmaM * 128) + k * 32),
                                    smem_desc_k_sw128(b_smem + c * (kPairHalfN * 128) + k * 32), idesc, (c | k) ? 1u : 0u);
                }
                tc_commit_pair(&bar_empty[s]);
                tc_commit_pair(&bar_tfull[a]);
            }
        }
    } else {
        // ===== epilogue (both CTAs, own rows): 8 warps = 4 TMEM lane quarters x 2 column halves of 64 =====
        const int q = warp & 3, h = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int R1 = 0, R2 = 0, Ridx = -1;
        for (int t = 0; t < ntiles; ++t) {
            const int a = t & 1;
            mbar_wait_or_trap(&bar_tfull[a], (t >> 1) & 1);
            tc_fence_after();
            const int col0 = (tile0 + t) * kTileN + h * 64;
            const int valid = nb - col0;
            uint32_t r[32];
            tmem_ld64_pack16(tmem_base + ((uint32_t)(q * 32) << 16) + a * kTileN + h * 64, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&bar_tempty[a], 0u);
            top2_tile_update(r, t, col0, valid, R1, R2, Ridx);
        }
        top2_merge_halves_and_store(h, row, row0 + row < na, R1, R2, Ridx, s_r1, s_r2, s_ri, o_idx + row0 + row, o_b1 + row0 + row, o_b2 + row0 + row);
    }
    tc_fence_before();
    cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still be signalled or read
    if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

// ---- debug: one tile, the raw dot products ----------------------------------------------------------------------
__global__ void __launch_bounds__(128) mma_dot_tile_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                           int32_t* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // SWIZZLE_128B blocks must be 1024-byte aligned
    uint8_t* sb = sa + kATileBytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar_load, kATileBytes + kBTileBytes);
        tma_load_2d(sa, &map_a, 0, 0, &bar_load);
        tma_load_2d(sa + kMmaM * 128, &map_a, 128, 0, &bar_load);
        tma_load_2d(sb, &map_b, 0, 0, &bar_load);
        tma_load_2d(sb + kMmaN * 128, &map_b, 128, 0, &bar_load);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) {
        mbar_wait_or_trap(&bar_load, 0);
        tc_fence_after();
        issue_tile_mmas(smem_u32(sa), smem_u32(sb), tmem_base);
        tc_commit(&bar_mma);
    }
    mbar_wait_or_trap(&bar_mma, 0);
    tc_fence_after();
    for (int chunk = 0; chunk < kMmaN / 32; ++chunk) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + chunk * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) out[(size_t)(warp * 32 + lane) * kMmaN + chunk * 32 + i] = (int32_t)r[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// host: 2-D tensor map over an expanded descriptor array [rows][256 bytes], box = 128 bytes x box_rows, SWIZZLE_128B
static int encode_expanded_map(CUtensorMap* out, const void* base, long long rows, int box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::atomic<void*> cached{nullptr};  // resolved once; racing threads resolve the same pointer
    void* p = cached.load();
    if (!p) {
        cudaDriverEntryPointQueryResult q;
        ORB_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return ORB_ECUDA; }
        cached.store(p);
    }
    const cuuint64_t dims[2] = {256u, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {256u};
    const cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = ((EncodeFn)p)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (expanded descriptors) failed with CUresult %d", (int)r); return ORB_ECUDA; }
    return ORB_OK;
}

}  // namespace orb

namespace orb {

// Scratch for the expanded operands: one pair of buffers per (device, stream), grown on demand and kept until the owner of
// the stream calls orbm_release_scratch. Calls on the same stream are ordered, so they share buffers; calls on different
// streams (two agents' frontends on one GPU) must not. The table lock covers the lookup only; growth (stream-ordered
// allocation, so no host or device-wide synchronisation) and tensor-map encoding happen under the entry's own lock.
struct MmaScratch {
    std::mutex mu;
    uint8_t* a = nullptr;
    uint8_t* b = nullptr;
    uint8_t* part = nullptr;           // partial results of a split candidate scan
    size_t cap_a = 0, cap_b = 0, cap_part = 0;
    CUtensorMap map_a, map_b, map_bh;  // cached: valid for (a, rows_a) / (b, rows_b); map_bh = b in boxes of half a tile (CTA pairs)
    size_t map_rows_a = 0, map_rows_b = 0;
};
using MmaScratchTable = std::map<std::pair<int, cudaStream_t>, std::shared_ptr<MmaScratch>>;
static std::mutex g_mma_mutex;
static MmaScratchTable& mma_scratch_table() {
    static MmaScratchTable* t = new MmaScratchTable;  // never destroyed: thread-exit hooks may run after static destructors
    return *t;
}

static int grow_stream_buffer(uint8_t** buf, size_t* cap, size_t bytes, cudaStream_t st) {
    if (bytes <= *cap) return ORB_OK;
    if (*buf) ORB_CUDA_TRY(cudaFreeAsync(*buf, st));  // ordered after the earlier calls on this stream that still read it
    *buf = nullptr; *cap = 0;
    const size_t want = bytes + bytes / 4;
    ORB_CUDA_TRY(cudaMallocAsync((void**)buf, want, st));
    *cap = want;
    return ORB_OK;
}

int release_mma_scratch(int device, cudaStream_t st) {
    std::shared_ptr<MmaScratch> sc;
    {
        std::lock_guard<std::mutex> lock(g_mma_mutex);
        MmaScratchTable& t = mma_scratch_table();
        auto it = t.find(std::make_pair(device, st));
        if (it == t.end()) return ORB_OK;
        sc = it->second;
        t.erase(it);
    }
    std::lock_guard<std::mutex> lock(sc->mu);
    if (sc->a) ORB_CUDA_TRY(cudaFreeAsync(sc->a, st));
    if (sc->b) ORB_CUDA_TRY(cudaFreeAsync(sc->b, st));
    if (sc->part) ORB_CUDA_TRY(cudaFreeAsync(sc->part, st));
    sc->a = sc->b = sc->part = nullptr;
    return ORB_OK;
}

// ---- candidate-range split ---------------------------------------------------------------------------------------
// Every CTA of a launch does the same amount of work, so a launch runs in waves: 1563 query tiles on 296 CTA slots are 5.28
// waves that cost 6. When that loses more than a few percent (and the database is long enough to keep every CTA's pipeline
// and pruning warm) the candidate tiles are dealt to gridDim.z CTAs per query tile, each leaving a partial (index, best,
// second); this kernel folds the partials in candidate order with the reference's update rule (strict '<' for the best,
// src/ORBmatcher.cc:589-598), so the result is the one-pass result.
constexpr int kMaxSplit = 4, kMinTilesPerSplit = 64;
__global__ void __launch_bounds__(256) top2_merge_splits_kernel(const int* __restrict__ part_idx, const int* __restrict__ part_b1,
                                                                const int* __restrict__ part_b2, long long split_stride, int nsplit,
                                                                const int* __restrict__ d_nA, int nA_max, const int* __restrict__ d_pairs,
                                                                int out_stride, long long total, int* __restrict__ out_idx,
                                                                int* __restrict__ out_b1, int* __restrict__ out_b2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int p = (int)(i / out_stride), r = (int)(i - (long long)p * out_stride);
    const int na = d_nA ? min(d_nA[d_pairs ? d_pairs[2 * p] : p], nA_max) : nA_max;
    if (r >= na) return;
    int idx = -1, m1 = 256, m2 = 256;
    for (int z = 0; z < nsplit; ++z) {
        const int d1 = part_b1[z * split_stride + i], d2 = part_b2[z * split_stride + i];
        if (d1 < m1) { m2 = min(m1, d2); m1 = d1; idx = part_idx[z * split_stride + i]; }
        else m2 = min(m2, d1);
    }
    out_idx[i] = idx; out_b1[i] = m1; out_b2[i] = m2;
}

// how many ways to split the candidate scan of `ctas` equal CTAs over `slots` resident ones: the smallest count within 3 %
// of the best wave efficiency
static int choose_split(long long ctas, int candidate_tiles, int slots) {
    int best = 1;
    double best_eff = 0.0;
    for (int n = 1; n <= kMaxSplit; ++n) {
        if (n > 1 && candidate_tiles / n < kMinTilesPerSplit) break;
        const long long units = ctas * n, waves = (units + slots - 1) / slots;
        const double eff = (double)units / (double)(waves * slots);
        if (eff > best_eff + 0.03) { best_eff = eff; best = n; }
    }
    return best;
}

// variant: 0 = automatic, 1 = one CTA per query tile, 2 = CTA pairs (cta_group::2). Measured on B200 with the candidate split in
// place (tools/exp_knn2_pair.py, profiles/r02_experiments.md): the pair kernel needs 3/4 of the shared-memory operand
// bandwidth (tc wavefronts 51 % against 68 %) but the tensor pipe is equally busy in both (72 %), and the one-CTA kernel is
// as fast or faster at every size - so the automatic choice is the one-CTA kernel and pairs run only when asked for.
static bool use_pair_kernel(int variant, int /*nA_max*/, int /*nB_max*/) { return variant == 2; }

static int set_mma_kernel_attributes(int device) {
    static std::atomic<bool> attr_set[64];  // per device, once: the attribute calls are not free on the launch path
    if (device >= 64 || !attr_set[device].load()) {
        ORB_CUDA_TRY(cudaFuncSetAttribute(knn2_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemBytes));
        ORB_CUDA_TRY(cudaFuncSetAttribute(knn2_mma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
        if (device < 64) attr_set[device].store(true);
    }
    return ORB_OK;
}

int launch_knn2_mma(const uint8_t* dA, const int* d_nA, int nA_max, int strideA_rows, const uint8_t* dB, const int* d_nB, int nB_max,
                    int strideB_rows, const int* d_pairs, int pairs, int out_stride, int* d_idx, int* d_b1, int* d_b2, cudaStream_t st,
                    int variant) {
    if (pairs <= 0 || nA_max <= 0) return ORB_OK;
    ORB_REQUIRE(pairs <= 65535, "more than 65535 set pairs in one call");
    int device = 0;
    ORB_CUDA_TRY(cudaGetDevice(&device));
    const size_t rows_a = (size_t)pairs * strideA_rows, rows_b = (size_t)pairs * strideB_rows;
    ORB_REQUIRE(rows_a < (1ull << 31) && rows_b < (1ull << 31), "too many descriptor rows for the tensor-core matcher");
    std::shared_ptr<MmaScratch> entry;
    {
        std::lock_guard<std::mutex> lock(g_mma_mutex);
        std::shared_ptr<MmaScratch>& slot = mma_scratch_table()[std::make_pair(device, st)];
        if (!slot) slot = std::make_shared<MmaScratch>();
        entry = slot;
    }
    std::lock_guard<std::mutex> entry_lock(entry->mu);
    MmaScratch& sc = *entry;
    int rc;
    const uint8_t *old_a = sc.a, *old_b = sc.b;
    if ((rc = grow_stream_buffer(&sc.a, &sc.cap_a, rows_a * 256, st)) != ORB_OK) return rc;
    if ((rc = grow_stream_buffer(&sc.b, &sc.cap_b, rows_b * 256, st)) != ORB_OK) return rc;
    if (sc.a != old_a || sc.map_rows_a != rows_a) {
        sc.map_rows_a = 0;
        if ((rc = encode_expanded_map(&sc.map_a, sc.a, (long long)rows_a, kMmaM)) != ORB_OK) return rc;
        sc.map_rows_a = rows_a;
    }
    if (sc.b != old_b || sc.map_rows_b != rows_b) {
        sc.map_rows_b = 0;
        if ((rc = encode_expanded_map(&sc.map_b, sc.b, (long long)rows_b, kTileN)) != ORB_OK) return rc;
        if ((rc = encode_expanded_map(&sc.map_bh, sc.b, (long long)rows_b, kPairHalfN)) != ORB_OK) return rc;
        sc.map_rows_b = rows_b;
    }
    const bool pair_kernel = use_pair_kernel(variant, nA_max, nB_max);
    const CUtensorMap ma = sc.map_a, mb = pair_kernel ? sc.map_bh : sc.map_b;
    if ((rc = set_mma_kernel_attributes(device)) != ORB_OK) return rc;
    const int qtiles = pair_kernel ? 2 * ceil_div(nA_max, 2 * kMmaM) : ceil_div(nA_max, kMmaM);
    const int nsplit = choose_split((long long)qtiles * pairs, ceil_div(nB_max, kTileN), 2 * kNumSMs);
    const long long total = (long long)pairs * out_stride;
    int *k_idx = d_idx, *k_b1 = d_b1, *k_b2 = d_b2;
    if (nsplit > 1) {
        if ((rc = grow_stream_buffer(&sc.part, &sc.cap_part, (size_t)3 * nsplit * total * sizeof(int), st)) != ORB_OK) return rc;
        k_idx = reinterpret_cast<int*>(sc.part);
        k_b1 = k_idx + (size_t)nsplit * total;
        k_b2 = k_b1 + (size_t)nsplit * total;
    }
    // expansion and matcher back to back in the stream (nothing in between: the matcher is a programmatic dependent)
    const int words = 8 * (nA_max > nB_max ? nA_max : nB_max);
    expand_pairs_kernel<<<dim3(ceil_div(words, 256), pairs, 2), 256, 0, st>>>((const uint32_t*)dA, d_nA, nA_max, strideA_rows, (const uint32_t*)dB,
                                                                             d_nB, nB_max, strideB_rows, d_pairs, (uint4*)sc.a, (uint4*)sc.b);
    if (pair_kernel)
        knn2_mma_pair_kernel<<<dim3(qtiles, pairs, nsplit), kMmaThreads, kPairSmemBytes, st>>>(
            ma, mb, d_nA, nA_max, strideA_rows, d_nB, nB_max, strideB_rows, d_pairs, out_stride, k_idx, k_b1, k_b2, total);
    else  // directly behind the expansion kernel in the stream: programmatic dependent launch
        ORB_CUDA_TRY(launch_pdl(knn2_mma_kernel, dim3(qtiles, pairs, nsplit), dim3(kMmaThreads), (size_t)kMmaSmemBytes, st, ma, mb, d_nA, nA_max,
                                strideA_rows, d_nB, nB_max, strideB_rows, d_pairs, out_stride, k_idx, k_b1, k_b2, total));
    if (nsplit > 1) {
        top2_merge_splits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(k_idx, k_b1, k_b2, total, nsplit, d_nA, nA_max, d_pairs, out_stride,
                                                                                 total, d_idx, d_b1, d_b2);
        count_launch();
    }
    count_launch(2);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb

namespace orb {
// ---- building blocks for the cross-map context (xmap.cu): operands already expanded by the caller --------------------
int mma_encode_operand_map(CUtensorMap* out, const void* base, long long rows, bool query_side) {
    return encode_expanded_map(out, base, rows, query_side ? kMmaM : kTileN);
}
// the database operand in boxes of half a candidate tile: what each CTA of a pair stages (knn2_mma_pair_kernel)
int mma_encode_operand_map_half(CUtensorMap* out, const void* base, long long rows) { return encode_expanded_map(out, base, rows, kPairHalfN); }
// rows [0, n_rows) of a packed 32-byte descriptor array (which may live in a PEER GPU's memory: the loads then travel over
// NVLink) -> 256 int8 per row, +-64 (query operand) or +-1 (database operand), into local memory
__global__ void __launch_bounds__(256) expand_rows_kernel(const uint32_t* __restrict__ src, int n_words, uint32_t neg, uint32_t flip,
                                                          uint4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint32_t w = src[i];
    uint32_t o[8];
#pragma unroll
    for (int nib = 0; nib < 8; ++nib) {
        const uint32_t x = (w >> (4 * nib)) & 0xFu;
        const uint32_t t = (x * 0x00204081u) & 0x01010101u;
        o[nib] = neg ^ (t * flip);
    }
    out[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
    out[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
}
int mma_expand_rows(const uint8_t* packed, int n_rows, bool query_side, uint8_t* out, cudaStream_t st) {
    if (n_rows <= 0) return ORB_OK;
    expand_rows_kernel<<<ceil_div(n_rows * 8, 256), 256, 0, st>>>((const uint32_t*)packed, n_rows * 8, query_side ? 0xC0C0C0C0u : 0xFFFFFFFFu,
                                                                  query_side ? 0x80u : 0xFEu, (uint4*)out);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}
// Loads the kernels of this file on the current device NOW. With CUDA's lazy module loading the first launch of a kernel
// loads its code, which can synchronise with work already running on that device; a rank whose stream is parked in a flag
// wait (xmap.cu) while the host is still enqueueing that rank's first matcher launch would then deadlock.
int mma_preload_kernels() {
    cudaFuncAttributes fa;
    ORB_CUDA_TRY(cudaFuncGetAttributes(&fa, expand_rows_kernel));
    ORB_CUDA_TRY(cudaFuncGetAttributes(&fa, knn2_mma_kernel));
    ORB_CUDA_TRY(cudaFuncGetAttributes(&fa, knn2_mma_pair_kernel));
    ORB_CUDA_TRY(cudaFuncGetAttributes(&fa, top2_merge_splits_kernel));
    int device = 0;
    ORB_CUDA_TRY(cudaGetDevice(&device));
    return set_mma_kernel_attributes(device);
}
// one query block (na rows, expanded, starting at row 0 of map_a) against one database (nb rows of map_b); the three
// output arrays may point into a peer GPU's memory
// `part`: scratch of mma_partial_ints(rows) ints for the partial results of a split candidate scan
size_t mma_partial_ints(size_t rows) { return (size_t)3 * kMaxSplit * rows; }
int mma_launch_preexpanded(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_bh, int na, int nb, int* d_idx, int* d_b1,
                           int* d_b2, int* part, cudaStream_t st) {
    if (na <= 0) return ORB_OK;
    int device = 0, rc;
    ORB_CUDA_TRY(cudaGetDevice(&device));
    if ((rc = set_mma_kernel_attributes(device)) != ORB_OK) return rc;
    const bool pair_kernel = use_pair_kernel(0, na, nb);
    const int qtiles = pair_kernel ? 2 * ceil_div(na, 2 * kMmaM) : ceil_div(na, kMmaM);
    const int nsplit = part ? choose_split(qtiles, ceil_div(nb, kTileN), 2 * kNumSMs) : 1;
    int *k_idx = d_idx, *k_b1 = d_b1, *k_b2 = d_b2;
    if (nsplit > 1) { k_idx = part; k_b1 = part + (size_t)nsplit * na; k_b2 = k_b1 + (size_t)nsplit * na; }
    if (pair_kernel)
        knn2_mma_pair_kernel<<<dim3(qtiles, 1, nsplit), kMmaThreads, kPairSmemBytes, st>>>(map_a, map_bh, nullptr, na, na, nullptr, nb, nb, nullptr, na,
                                                                                         k_idx, k_b1, k_b2, (long long)na);
    else
        knn2_mma_kernel<<<dim3(qtiles, 1, nsplit), kMmaThreads, kMmaSmemBytes, st>>>(map_a, map_b, nullptr, na, na, nullptr, nb, nb, nullptr, na, k_idx,
                                                                                   k_b1, k_b2, (long long)na);
    count_launch();
    if (nsplit > 1) {
        top2_merge_splits_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(k_idx, k_b1, k_b2, (long long)na, nsplit, nullptr, na, nullptr, na,
                                                                              (long long)na, d_idx, d_b1, d_b2);
        count_launch();
    }
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}
}  // namespace orb

// Device-call users own their streams: before destroying one (or to give the memory back) they release the expanded-operand
// scratch the brute-force matcher keeps for it. Stream-ordered, no synchronisation.
extern "C" int orbm_release_scratch(void* stream) {
    using namespace orb;
    int device = 0;
    ORB_CUDA_TRY(cudaGetDevice(&device));
    return release_mma_scratch(device, (cudaStream_t)stream);
}

// Debug: how many CTA pairs of knn2_mma_pair_kernel the current device holds at once, and one-CTA blocks per SM of knn2_mma_kernel.
extern "C" int orbm_debug_mma_occupancy(int* pair_clusters, int* single_blocks_per_sm) {
    using namespace orb;
    ORB_REQUIRE(pair_clusters && single_blocks_per_sm, "null pointer");
    int device = 0, rc;
    ORB_CUDA_TRY(cudaGetDevice(&device));
    if ((rc = set_mma_kernel_attributes(device)) != ORB_OK) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * kNumSMs, 1, 1);
    cfg.blockDim = dim3(kMmaThreads, 1, 1);
    cfg.dynamicSmemBytes = kPairSmemBytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    ORB_CUDA_TRY(cudaOccupancyMaxActiveClusters(pair_clusters, knn2_mma_pair_kernel, &cfg));
    ORB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(single_blocks_per_sm, knn2_mma_kernel, kMmaThreads, kMmaSmemBytes));
    return ORB_OK;
}

// Debug / experiment entry: the tensor-core matcher on one pair of device arrays.
extern "C" int orbm_knn2_mma_device(const uint8_t* dA, int nA, const uint8_t* dB, int nB, int32_t* d_idx, int32_t* d_best, int32_t* d_second,
                                    void* stream) {
    using namespace orb;
    ORB_REQUIRE(nA >= 0 && nB >= 0, "negative row count");
    ORB_REQUIRE(nA == 0 || (dA && d_idx && d_best && d_second), "null pointer");
    return launch_knn2_mma(dA, nullptr, nA, nA, dB, nullptr, nB, nB, nullptr, 1, nA, d_idx, d_best, d_second, (cudaStream_t)stream, 0);
}

// Debug tap: dot products (+-1 encoding) of 128 x 256 descriptors through expand + TMA + tcgen05.mma + tcgen05.ld.
extern "C" int orbm_debug_mma_dot(int device, const uint8_t* A128, const uint8_t* B256, int32_t* out) {
    using namespace orb;
    ORB_REQUIRE(A128 && B256 && out, "null pointer");
    ORB_CUDA_TRY(cudaSetDevice(device));
    uint8_t *d_bits = nullptr, *d_exp = nullptr;
    int32_t* d_out = nullptr;
    ORB_CUDA_TRY(cudaMalloc(&d_bits, (128 + 256) * 32));
    ORB_CUDA_TRY(cudaMalloc(&d_exp, (128 + 256) * 256));
    ORB_CUDA_TRY(cudaMalloc(&d_out, 128 * 256 * 4));
    ORB_CUDA_TRY(cudaMemcpy(d_bits, A128, 128 * 32, cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(d_bits + 128 * 32, B256, 256 * 32, cudaMemcpyHostToDevice));
    expand_pm1_kernel<<<ceil_div((128 + 256) * 8, 256), 256>>>((const uint32_t*)d_bits, (128 + 256) * 8, (uint4*)d_exp);
    CUtensorMap ma, mb;
    int rc = encode_expanded_map(&ma, d_exp, 128, 128);
    if (rc == ORB_OK) rc = encode_expanded_map(&mb, d_exp + 128 * 256, 256, 256);
    if (rc == ORB_OK) {
        ORB_CUDA_TRY(cudaFuncSetAttribute(mma_dot_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kATileBytes + kBTileBytes + 1024));
        mma_dot_tile_kernel<<<1, 128, kATileBytes + kBTileBytes + 1024>>>(ma, mb, d_out);
        count_launch(2);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { set_error("mma_dot_tile_kernel failed: %s", cudaGetErrorString(e)); rc = ORB_ECUDA; }
        else ORB_CUDA_TRY(cudaMemcpy(out, d_out, 128 * 256 * 4, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_bits); cudaFree(d_exp); cudaFree(d_out);
    return rc;
}
