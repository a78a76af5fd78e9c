// Stand-alone check of the 3-D u8 TMA box load used by the stencil kernels.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

struct Maps { CUtensorMap m[16]; };

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ Maps maps, int level, int x, int y, int z, int bw, int bh, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bw * bh) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)),
                     "l"(reinterpret_cast<uint64_t>(&maps.m[level])), "r"(x), "r"(y), "r"(z), "r"(s32(&bar))
                     : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    const int w = 752, h = 480, frames = 4, pitch = 768, bw = 48, bh = 38;
    std::vector<uint8_t> img((size_t)pitch * h * frames);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d, *o;
    cudaMalloc(&d, img.size()); cudaMalloc(&o, 65536);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry %p q=%d\n", p, (int)q);
    Maps maps{};
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    const cuuint32_t box[3] = {bw, bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)p)(&maps.m[3], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode r=%d\n", (int)r);
    const int ax = argc > 1 ? atoi(argv[1]) : 16, ay = argc > 2 ? atoi(argv[2]) : 16, az = argc > 3 ? atoi(argv[3]) : 0;
    for (int t = 0; t < 1; ++t) {
        const int x = ax, y = ay, z = az;
        k<<<1, 128, bw * bh>>>(maps, 3, x, y, z, bw, bh, o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("case %d: %s\n", t, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> got(bw * bh);
        cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
            const int gx = x + xx, gy = y + yy;
            const uint8_t ref = (gx < 0 || gx >= w || gy < 0 || gy >= h) ? 0 : img[(size_t)z * pitch * h + (size_t)gy * pitch + gx];
            bad += got[yy * bw + xx] != ref;
        }
        printf("  mismatches: %d\n", bad);
    }
    return 0;
}
