// K1: pyramid level resize, K2: 7x7 Gaussian blur. Integer fixed point, bit-exact with OpenCV:
//   cv::resize(INTER_LINEAR, 8U)   called at /root/reference/src/ORBextractor.cc:1120
//   cv::GaussianBlur(7x7, sigma 2) called at src/ORBextractor.cc:1085-1086
// Arithmetic pinned in SURVEY.md Appendix A.1 / A.2 and oracle/cvprim.h.
//
// Both are streaming stencils: HBM roofline. Algorithmic bytes per frame:
//   resize level l : read w_{l-1}*h_{l-1}, write w_l*h_l
//   blur           : read + write sum_l w_l*h_l
#include "extract_kernels.cuh"

namespace orb {

// ------------------------------------------------------------------------------------------------
// resize: each thread produces 4 horizontally adjacent pixels of one output row.
// grid (ceil(w/128), ceil(h/8), frames), block (32, 8)
__global__ void __launch_bounds__(256)
resize_kernel(const Geometry* __restrict__ g, const LinTap* __restrict__ taps, FrameSet fs, uint8_t* __restrict__ pyr, int level) {
    const LevelGeom& L = g->lv[level];
    const LevelGeom& S = g->lv[level - 1];
    const int frame = blockIdx.z;
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x0 >= L.w || y >= L.h) return;
    int spitch;
    const uint8_t* src = level_ptr(*g, fs, pyr, frame, level - 1, &spitch);
    uint8_t* dst = pyr + (size_t)frame * g->pyr_bytes + L.img_off;

    const LinTap ty = taps[L.tab_y_off + y];
    const uint8_t* r0 = src + (size_t)ty.ofs * spitch;
    const uint8_t* r1 = src + (size_t)min(ty.ofs + 1, S.h - 1) * spitch;
    const int b0 = ty.c0, b1 = ty.c1;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x0 + k;
        if (x < L.w) {
            const LinTap tx = taps[L.tab_x_off + x];
            const int sx0 = tx.ofs, sx1 = min(tx.ofs + 1, S.w - 1);
            const int h0 = r0[sx0] * tx.c0 + r0[sx1] * tx.c1;
            const int h1 = r1[sx0] * tx.c0 + r1[sx1] * tx.c1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            packed |= (uint32_t)(v & 0xff) << (8 * k);
        }
    }
    // rows are 128-byte pitched and x0 is a multiple of 4: aligned 32-bit store (pitch slack absorbs the tail)
    *reinterpret_cast<uint32_t*>(dst + (size_t)y * L.pitch + x0) = packed;
}

int launch_resize_level(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int level, int n, cudaStream_t st) {
    const LevelGeom& L = hg.lv[level];
    dim3 grid(ceil_div(L.w, 128), ceil_div(L.h, 8), n);
    resize_kernel<<<grid, dim3(32, 8), 0, st>>>(db.geom, db.taps, fs, db.pyr, level);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// ------------------------------------------------------------------------------------------------
// blur: separable {18,34,48,56,48,34,18}, single rounding (sum + 2^15) >> 16, BORDER_REFLECT_101.
// One block = one 128 x 16 output tile of one level of one frame; the (128+6) x (16+6) input halo
// is staged in shared memory, the horizontal pass keeps 16-bit sums in shared memory.
constexpr int kBlurTW = 128, kBlurTH = 16;

__device__ __forceinline__ int reflect101(int p, int n) {
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * n - 2 - p : p;
    return max(0, min(p, n - 1));  // only out-of-tile lanes of partial tiles ever hit the clamp
}

__global__ void __launch_bounds__(256)
blur_kernel(const Geometry* __restrict__ g, const BlurTile* __restrict__ tiles, FrameSet fs, const uint8_t* __restrict__ pyr,
            uint8_t* __restrict__ blur) {
    __shared__ uint8_t in[kBlurTH + 6][kBlurTW + 8];
    __shared__ uint16_t hs[kBlurTH + 6][kBlurTW];
    const BlurTile t = tiles[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelGeom& L = g->lv[t.level];
    int spitch;
    const uint8_t* src = level_ptr(*g, fs, pyr, frame, t.level, &spitch);
    uint8_t* dst = blur + (size_t)frame * g->blur_bytes + L.blur_off;
    const int X0 = t.tx * kBlurTW, Y0 = t.ty * kBlurTH;
    const int tid = threadIdx.x;

    for (int i = tid; i < (kBlurTH + 6) * (kBlurTW + 6); i += 256) {
        const int r = i / (kBlurTW + 6), c = i - r * (kBlurTW + 6);
        const int sy = reflect101(Y0 + r - 3, L.h), sx = reflect101(X0 + c - 3, L.w);
        in[r][c] = src[(size_t)sy * spitch + sx];
    }
    __syncthreads();
    for (int i = tid; i < (kBlurTH + 6) * kBlurTW; i += 256) {
        const int r = i / kBlurTW, c = i - r * kBlurTW;
        const uint8_t* p = &in[r][c];
        hs[r][c] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    // vertical pass: each thread 4 adjacent pixels of a row, 32-bit store
    for (int i = tid; i < kBlurTH * (kBlurTW / 4); i += 256) {
        const int r = i / (kBlurTW / 4), c4 = (i - r * (kBlurTW / 4)) * 4;
        const int y = Y0 + r, x = X0 + c4;
        if (y >= L.h || x >= L.w) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c4 + k;
            const uint32_t s = 18u * (hs[r][c] + hs[r + 6][c]) + 34u * (hs[r + 1][c] + hs[r + 5][c]) +
                               48u * (hs[r + 2][c] + hs[r + 4][c]) + 56u * hs[r + 3][c] + 32768u;
            packed |= (s >> 16) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * L.pitch + x) = packed;
    }
}

int launch_blur(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int n, cudaStream_t st) {
    blur_kernel<<<dim3(hg.ntiles, n), 256, 0, st>>>(db.geom, db.tiles, fs, db.pyr, db.blur);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
