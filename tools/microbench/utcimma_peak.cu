// Micro-benchmark: dense int8 tensor-core peak of this GPU as the Hamming matcher can see it - a bare issue loop of
// tcgen05.mma.cta_group::1.kind::i8 (SASS: UTCIMMA) with both operands resident in shared memory (SWIZZLE_128B K-major
// tiles, the layout of csrc/hamming_mma.cu) and the accumulator in TMEM; no loads, no epilogue. One thread per CTA issues
// `iters` x 4 MMAs (M = 128, N = 128 or 256, K = 32 each) and commits once; the CTA waits on the mbarrier.
//   int8 ops = CTAs x iters x 4 x 2 x 128 x N x 32
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o utcimma_peak utcimma_peak.cu
// Output: one JSON line per configuration (N, CTAs per SM), the best one is the denominator of bench.py's tensor fraction.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1u << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1u << 46;
    d |= (uint64_t)2u << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// COMMITS: extra tcgen05.commit instructions (to a barrier nobody waits on) after every 8 MMAs - what a real pipeline does to
// hand shared-memory slots and accumulators over; ACCS: accumulators used in rotation (1 = always the same TMEM columns)
template <int N, int COMMITS = 0, int ACCS = 1>
__global__ void __launch_bounds__(128) peak_kernel(int iters) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar, bar_dummy;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sb = sa + 128 * 128;
    for (int i = threadIdx.x; i < (128 + N) * 128 / 4; i += blockDim.x) ((uint32_t*)sa)[i] = 0x01FF01FFu * (i | 1);
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)(N * ACCS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_dummy)), "r"(1u << 20));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_i8(128, N);
        const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = desc_sw128(a0 + k * 32), db = desc_sw128(b0 + k * 32);
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem + (uint32_t)(((it >> 1) % ACCS) * N)),
                    "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(it | k)), "r"(0u)
                    : "memory");
            }
            if (COMMITS && (it & 1)) {
#pragma unroll
                for (int c = 0; c < COMMITS; ++c)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_dummy)) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everybody waits for the MMAs
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(&bar)), "r"(0u)
            : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(N * ACCS)) : "memory");
}

template <int N, int COMMITS = 0, int ACCS = 1>
static double run(int ctas_per_sm, int iters, int sms) {
    const size_t smem = (128 + N) * 128 + 1024;
    cudaFuncSetAttribute(peak_kernel<N, COMMITS, ACCS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = sms * ctas_per_sm;
    peak_kernel<N, COMMITS, ACCS><<<grid, 128, smem>>>(64);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        peak_kernel<N, COMMITS, ACCS><<<grid, 128, smem>>>(iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(e)); return -1; }
    const double ops = (double)grid * iters * 4 * 2.0 * 128 * N * 32;
    const double tops = ops / (best * 1e-3) / 1e12;
    printf("{\"kernel\": \"tcgen05.mma kind::i8 128x%dx32 issue loop\", \"commits_per_8_mma\": %d, \"accumulators\": %d, \"ctas_per_sm\": %d, \"iters\": %d, "
           "\"ms\": %.4f, \"int8_tops\": %.1f}\n", N, COMMITS, ACCS, ctas_per_sm, iters, best, tops);
    return tops;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double best = 0;
    for (int c = 1; c <= 2; ++c) {
        double t = run<128>(c, 20000, sms); if (t > best) best = t;
        t = run<256>(c, 10000, sms); if (t > best) best = t;
    }
    // what the hand-over points of a real pipeline cost: commits after every 8 MMAs, rotating accumulators
    run<128, 1, 1>(1, 20000, sms); run<128, 2, 1>(1, 20000, sms); run<128, 2, 2>(1, 20000, sms); run<128, 2, 2>(2, 20000, sms);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"int8_tops_measured\": %.1f}\n", p.name, sms, best);
    return best > 0 ? 0 : 1;
}
