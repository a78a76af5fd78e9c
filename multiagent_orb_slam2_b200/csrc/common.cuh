// Shared helpers of the B200 ORB frontend kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/orb_b200.h"
#include "../../include/orb_b200_debug.h"

namespace orb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);  // kernels launched by this library (bench.py's gpu_launches)
extern "C" long long orb_launch_count(void);

#define ORB_CUDA_TRY(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            orb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return ORB_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

#define ORB_REQUIRE(cond, msg)                          \
    do {                                                \
        if (!(cond)) {                                  \
            orb::set_error("invalid argument: %s", msg); \
            return ORB_EINVAL;                          \
        }                                               \
    } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- host-buffer wrappers: H2D, kernel, D2H, synchronise --------------------------------------
// ORBmatcher is constructed on the stack at 13 call sites and called from the Tracking, LocalMapping, LoopClosing and
// MapFusion threads concurrently, the vocabulary is shared by all of them: every calling thread keeps
// its own device arena (grow-only), pinned staging and stream, so a host-buffer call costs copies + launches, not a
// cudaMalloc / cudaFree pair per array on the default stream.
int release_mma_scratch(int device, cudaStream_t st);   // hamming_mma.cu: operand scratch keyed by (device, stream)
struct HostCallWorkspace {
    int device = -1;
    cudaStream_t st = nullptr;
    uint8_t* base = nullptr;
    size_t cap = 0, used = 0;
    uint8_t* pin = nullptr;   // pinned staging for the results: ONE device -> host copy per call instead of one blocking
    size_t pin_cap = 0;       // copy into pageable memory per output array (~15 us each)
    ~HostCallWorkspace() { release(); }
    int pinned(size_t bytes, uint8_t** out) {
        if (bytes > pin_cap) {
            if (pin) { cudaFreeHost(pin); pin = nullptr; pin_cap = 0; }
            ORB_CUDA_TRY(cudaMallocHost(&pin, bytes + bytes / 2 + 4096));
            pin_cap = bytes + bytes / 2 + 4096;
        }
        *out = pin;
        return ORB_OK;
    }
    void release() {
        if (pin) { cudaFreeHost(pin); pin = nullptr; pin_cap = 0; }
        if (device >= 0) {
            cudaSetDevice(device);
            if (st) release_mma_scratch(device, st);
            if (base) cudaFree(base);
            if (st) cudaStreamDestroy(st);
        }
        base = nullptr; st = nullptr; cap = 0; device = -1;
    }
    // make room for `bytes` on `dev`; pointers handed out before are invalid afterwards
    int begin(int dev, size_t bytes) {
        if (dev != device) {
            release();
            ORB_CUDA_TRY(cudaSetDevice(dev));
            ORB_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            device = dev;
        } else {
            ORB_CUDA_TRY(cudaSetDevice(dev));
        }
        if (bytes > cap) {
            if (base) { cudaStreamSynchronize(st); cudaFree(base); base = nullptr; cap = 0; }
            const size_t want = bytes + bytes / 2 + 4096;
            ORB_CUDA_TRY(cudaMalloc(&base, want));
            cap = want;
        }
        used = 0;
        return ORB_OK;
    }
    template <class T> T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + used);
        used += align_up(count * sizeof(T) + 1, 256);
        return p;
    }
    static size_t need(size_t bytes) { return align_up(bytes + 1, 256); }
};
HostCallWorkspace& host_call_workspace();   // the calling thread's (runtime.cu)

// Launch `kernel` as a programmatic dependent of the previous kernel in `st`: it may be scheduled while that kernel's last
// blocks still run and must execute griddepcontrol.wait before it touches anything the predecessor writes.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }            // no-ops in an ordinary launch
__device__ __forceinline__ void pdl_release_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

}  // namespace orb
