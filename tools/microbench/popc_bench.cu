// Micro-benchmark: integer-pipe throughput relevant to Hamming matching on sm_100a.
//   (a) POPC.b32 results / clk / SM     (b) LOP3 / clk / SM    (c) IADD3 / clk / SM
//   (d) 256-bit Hamming distance: plain 8 x POPC  vs  carry-save-adder tree (4 x POPC + LOP3s)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o popc_bench popc_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void pipe_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t acc = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = __popc(a[i]) + a[i];           // POPC + IADD (dependent chain x8 independent)
            if (MODE == 1) a[i] = (a[i] ^ acc) & (a[(i + 1) & 7] | it);  // LOP3
            if (MODE == 2) a[i] = a[i] + a[(i + 1) & 7] + it;    // IADD3
        }
        acc += a[0];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

template <int MODE>
__global__ void hamming_kernel(const uint4* __restrict__ db, int ndb, uint32_t* out) {
    __shared__ uint4 tile[512];
    const uint4 q0 = db[(threadIdx.x * 2) % (2 * ndb)], q1 = db[(threadIdx.x * 2 + 1) % (2 * ndb)];
    int best = 256, second = 256;
    for (int base = 0; base < ndb; base += 256) {
        __syncthreads();
        for (int i = threadIdx.x; i < 512; i += blockDim.x) tile[i] = db[base * 2 + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < 256; ++j) {
            const uint4 b0 = tile[2 * j], b1 = tile[2 * j + 1];
            const uint32_t x0 = q0.x ^ b0.x, x1 = q0.y ^ b0.y, x2 = q0.z ^ b0.z, x3 = q0.w ^ b0.w;
            const uint32_t x4 = q1.x ^ b1.x, x5 = q1.y ^ b1.y, x6 = q1.z ^ b1.z, x7 = q1.w ^ b1.w;
            int d;
            if (MODE == 0) {
                d = __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
            } else {
                uint32_t c0, s0, c1, s1, c2, s2, t0, d0;
                csa(c0, s0, x0, x1, x2);
                csa(c1, s1, x3, x4, x5);
                csa(c2, s2, s0, s1, x6);
                const uint32_t ones = s2 ^ x7, c3 = s2 & x7;
                csa(d0, t0, c0, c1, c2);
                const uint32_t twos = t0 ^ c3, d1 = t0 & c3;
                const uint32_t fours = d0 ^ d1, eights = d0 & d1;
                d = __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
            }
            second = min(second, max(d, best));
            best = min(best, d);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = best + second;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.0f MHz nominal\n", p.name, sms, clk_khz / 1e3);
    uint32_t* out;
    cudaMalloc(&out, 1 << 26);
    const int blocks = sms * 8, threads = 256;
    const char* names[3] = {"POPC(+IADD)", "LOP3", "IADD3"};
    for (int m = 0; m < 3; ++m) {
        float ms = 0;
        if (m == 0) ms = time_ms([&] { pipe_kernel<0><<<blocks, threads>>>(out, 12345); });
        if (m == 1) ms = time_ms([&] { pipe_kernel<1><<<blocks, threads>>>(out, 12345); });
        if (m == 2) ms = time_ms([&] { pipe_kernel<2><<<blocks, threads>>>(out, 12345); });
        const double ops = (double)blocks * threads * ITERS * 8;
        printf("%-12s %8.3f ms  %8.1f Gop/s  %6.1f op/clk/SM (at nominal clock)\n", names[m], ms, ops / ms / 1e6,
               ops / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    const int ndb = 8192;
    uint4* db;
    cudaMalloc(&db, ndb * 32);
    cudaMemset(db, 0x5a, ndb * 32);
    for (int m = 0; m < 2; ++m) {
        float ms = m == 0 ? time_ms([&] { hamming_kernel<0><<<blocks, threads>>>(db, ndb, out); })
                          : time_ms([&] { hamming_kernel<1><<<blocks, threads>>>(db, ndb, out); });
        const double cmps = (double)blocks * threads * ndb;
        printf("hamming256 %-6s %8.3f ms  %8.1f Gcmp/s  %6.2f cmp/clk/SM\n", m ? "CSA" : "8xPOPC", ms, cmps / ms / 1e6,
               cmps / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
