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
// resize: one block = one 128 x 32 output tile. The source window of the tile is staged in shared
// memory with aligned 32-bit loads; each thread keeps the column taps of its 4 output columns and
// the row taps of its 4 output rows in registers (fetched up front, so the kernel pays two global
// latencies - taps, then window - instead of three).
constexpr int kRsTW = 128, kRsTH = 32;
constexpr int kRsMaxWords = 72, kRsMaxRows = 68;  // source window budget: scale factors up to 2.0

__global__ void __launch_bounds__(256)
resize_kernel(const Geometry* __restrict__ g, const LinTap* __restrict__ taps, FrameSet fs, uint8_t* __restrict__ pyr, int level) {
    __shared__ __align__(16) uint32_t win[kRsMaxRows][kRsMaxWords];
    const LevelGeom& L = g->lv[level];
    const LevelGeom& S = g->lv[level - 1];
    const int frame = blockIdx.z;
    const int X0 = blockIdx.x * kRsTW, Y0 = blockIdx.y * kRsTH;
    const int X1 = min(X0 + kRsTW, L.w) - 1, Y1 = min(Y0 + kRsTH, L.h) - 1;  // last output column / row of the tile
    int spitch;
    const uint8_t* src = level_ptr(*g, fs, pyr, frame, level - 1, &spitch);
    uint8_t* dst = pyr + (size_t)frame * g->pyr_bytes + L.img_off;
    const LinTap* tx = taps + L.tab_x_off;
    const LinTap* ty = taps + L.tab_y_off;

    // all table reads of this thread, issued together
    const int q = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int x0 = X0 + 4 * q;
    LinTap tcol[4], trow[kRsTH / 8];
#pragma unroll
    for (int k = 0; k < 4; ++k) tcol[k] = tx[min(x0 + k, L.w - 1)];
#pragma unroll
    for (int rr = 0; rr < kRsTH / 8; ++rr) trow[rr] = ty[min(Y0 + rg + 8 * rr, L.h - 1)];
    // source window [sx_lo, sx_hi] x [sy_lo, sy_hi] (taps are monotone)
    const int sx_lo = tx[X0].ofs & ~3, sx_hi = min(tx[X1].ofs + 1, S.w - 1);
    const int sy_lo = ty[Y0].ofs, sy_hi = min(ty[Y1].ofs + 1, S.h - 1);
    const int nwords = (sx_hi - sx_lo) / 4 + 1, nrows = sy_hi - sy_lo + 1;
    const bool word_rows = ((spitch & 3) == 0) && (((uintptr_t)src & 3) == 0);
    {
        int r = threadIdx.x / nwords, cw = threadIdx.x - r * nwords;
        const int dr = 256 / nwords, dcw = 256 - dr * nwords;
        for (int i = threadIdx.x; i < nrows * nwords; i += 256) {
            const uint8_t* row = src + (size_t)(sy_lo + r) * spitch;
            const int x = sx_lo + 4 * cw;
            uint32_t v;
            if (word_rows && x + 3 < S.w) {
                v = __ldg(reinterpret_cast<const uint32_t*>(row + x));
            } else {
                v = 0;
                for (int b = 0; b < 4; ++b) v |= (uint32_t)row[min(x + b, S.w - 1)] << (8 * b);
            }
            win[r][cw] = v;
            cw += dcw; r += dr;
            if (cw >= nwords) { cw -= nwords; ++r; }
        }
    }
    __syncthreads();

    if (x0 >= L.w) return;
    int ofs0[4], ofs1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { ofs0[k] = tcol[k].ofs - sx_lo; ofs1[k] = min(tcol[k].ofs + 1, S.w - 1) - sx_lo; }
    const uint8_t* wb = reinterpret_cast<const uint8_t*>(&win[0][0]);
#pragma unroll
    for (int rr = 0; rr < kRsTH / 8; ++rr) {
        const int y = Y0 + rg + 8 * rr;
        if (y >= L.h) break;
        const LinTap t = trow[rr];
        const uint8_t* r0 = wb + (size_t)(t.ofs - sy_lo) * (kRsMaxWords * 4);
        const uint8_t* r1 = wb + (size_t)(min(t.ofs + 1, S.h - 1) - sy_lo) * (kRsMaxWords * 4);
        const int b0 = t.c0, b1 = t.c1;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h0 = r0[ofs0[k]] * tcol[k].c0 + r0[ofs1[k]] * tcol[k].c1;
            const int h1 = r1[ofs0[k]] * tcol[k].c0 + r1[ofs1[k]] * tcol[k].c1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            packed |= (uint32_t)(v & 0xff) << (8 * k);
        }
        // rows are 128-byte pitched and x0 is a multiple of 4: aligned 32-bit store (pitch slack absorbs the tail)
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * L.pitch + x0) = packed;
    }
}

int launch_resize_level(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int level, int n, cudaStream_t st) {
    const LevelGeom& L = hg.lv[level];
    dim3 grid(ceil_div(L.w, kRsTW), ceil_div(L.h, kRsTH), n);
    resize_kernel<<<grid, 256, 0, st>>>(db.geom, db.taps, fs, db.pyr, level);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// ------------------------------------------------------------------------------------------------
// blur: separable {18,34,48,56,48,34,18}, single rounding (sum + 2^15) >> 16, BORDER_REFLECT_101.
// One block = one 128 x 58 output tile of one level of one frame (input 136 x 64 incl. halo).
//   load : aligned 32-bit words of the input rows into shared memory (reflected bytes at the borders)
//   h    : 4 outputs per thread from 3 words: byte windows by funnel shift, 2 x IDP.4A per output
//          (4+3 taps, u8 x u8 -> u32, exact); rows are processed in vertical pairs and stored as
//          (row 2m | row 2m+1 << 16) so that
//   v    : one IDP.2A covers two vertical taps: 4 x IDP.2A per output, accumulator preloaded with
//          the rounding constant; 4 columns per thread, sliding window down the tile.
// ~14 instructions per pixel instead of ~120 for the straightforward byte-wise version.
constexpr int kBlurTW = 128, kBlurTH = 58, kBlurInRows = 64, kBlurInWords = 34;

__device__ __forceinline__ int reflect101(int p, int n) {
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * n - 2 - p : p;
    return max(0, min(p, n - 1));  // only out-of-tile lanes of partial tiles ever hit the clamp
}

__global__ void __launch_bounds__(256)
blur_kernel(const Geometry* __restrict__ g, const BlurTile* __restrict__ tiles, FrameSet fs, const uint8_t* __restrict__ pyr,
            uint8_t* __restrict__ blur) {
    __shared__ __align__(16) uint32_t in_w[kBlurInRows][kBlurInWords];
    __shared__ __align__(16) uint32_t hsv[kBlurInRows / 2][kBlurTW];
    const BlurTile t = tiles[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelGeom& L = g->lv[t.level];
    int spitch;
    const uint8_t* src = level_ptr(*g, fs, pyr, frame, t.level, &spitch);
    uint8_t* dst = blur + (size_t)frame * g->blur_bytes + L.blur_off;
    const int X0 = t.tx * kBlurTW, Y0 = t.ty * kBlurTH;
    const int tid = threadIdx.x;
    const bool word_rows = ((spitch & 3) == 0) && (((uintptr_t)src & 3) == 0);

    for (int i = tid; i < kBlurInRows * kBlurInWords; i += 256) {
        const int r = i / kBlurInWords, cw = i - r * kBlurInWords;
        const uint8_t* row = src + (size_t)reflect101(Y0 + r - 3, L.h) * spitch;
        const int x = X0 - 4 + 4 * cw;
        uint32_t v;
        if (word_rows && x >= 0 && x + 3 < L.w) {
            v = __ldg(reinterpret_cast<const uint32_t*>(row + x));
        } else {
            v = (uint32_t)row[reflect101(x, L.w)] | (uint32_t)row[reflect101(x + 1, L.w)] << 8 |
                (uint32_t)row[reflect101(x + 2, L.w)] << 16 | (uint32_t)row[reflect101(x + 3, L.w)] << 24;
        }
        in_w[r][cw] = v;
    }
    __syncthreads();

    // horizontal pass: item = (row pair, column quad)
    constexpr uint32_t kLo = 18u | 34u << 8 | 48u << 16 | 56u << 24, kHi = 48u | 34u << 8 | 18u << 16;
    for (int it = tid; it < (kBlurInRows / 2) * (kBlurTW / 4); it += 256) {
        const int pr = it >> 5, k = it & 31;
        uint32_t h[2][4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const uint32_t w0 = in_w[2 * pr + rr][k], w1 = in_w[2 * pr + rr][k + 1], w2 = in_w[2 * pr + rr][k + 2];
            // output j (column 4k+j) uses bytes j+1..j+4 and j+5..j+8 of the 12-byte string w0 w1 w2
            const uint32_t a1 = __funnelshift_r(w0, w1, 8), a2 = __funnelshift_r(w0, w1, 16), a3 = __funnelshift_r(w0, w1, 24);
            const uint32_t b1 = __funnelshift_r(w1, w2, 8), b2 = __funnelshift_r(w1, w2, 16), b3 = __funnelshift_r(w1, w2, 24);
            h[rr][0] = __dp4a(a1, kLo, __dp4a(b1, kHi, 0u));
            h[rr][1] = __dp4a(a2, kLo, __dp4a(b2, kHi, 0u));
            h[rr][2] = __dp4a(a3, kLo, __dp4a(b3, kHi, 0u));
            h[rr][3] = __dp4a(w1, kLo, __dp4a(w2, kHi, 0u));
        }
        uint4 o;
        o.x = h[0][0] | h[1][0] << 16; o.y = h[0][1] | h[1][1] << 16; o.z = h[0][2] | h[1][2] << 16; o.w = h[0][3] | h[1][3] << 16;
        *reinterpret_cast<uint4*>(&hsv[pr][4 * k]) = o;
    }
    __syncthreads();

    // vertical pass: thread = (column quad, segment of 4 output row pairs); output rows 2m, 2m+1 use
    // the pair rows m..m+3
    constexpr uint32_t e0 = 18u | 34u << 8, e1 = 48u | 56u << 8, e2 = 48u | 34u << 8, e3 = 18u;       // even row 2m
    constexpr uint32_t o0 = 18u << 8, o1 = 34u | 48u << 8, o2 = 56u | 48u << 8, o3 = 34u | 18u << 8;  // odd row 2m+1
    const int quad = tid & 31, seg = tid >> 5;
    const int x = X0 + 4 * quad;
    if (x < L.w) {
        const int m0 = seg * 4;
        uint4 p0 = *reinterpret_cast<const uint4*>(&hsv[m0][4 * quad]);
        uint4 p1 = *reinterpret_cast<const uint4*>(&hsv[m0 + 1][4 * quad]);
        uint4 p2 = *reinterpret_cast<const uint4*>(&hsv[m0 + 2][4 * quad]);
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
            const int m = m0 + mm;
            if (m >= kBlurTH / 2) break;
            const uint4 p3 = *reinterpret_cast<const uint4*>(&hsv[m + 3][4 * quad]);
            const uint32_t c0[4] = {p0.x, p0.y, p0.z, p0.w}, c1[4] = {p1.x, p1.y, p1.z, p1.w};
            const uint32_t c2[4] = {p2.x, p2.y, p2.z, p2.w}, c3[4] = {p3.x, p3.y, p3.z, p3.w};
            uint32_t ev = 0, od = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t se = __dp2a_lo(c0[j], e0, __dp2a_lo(c1[j], e1, __dp2a_lo(c2[j], e2, __dp2a_lo(c3[j], e3, 32768u))));
                const uint32_t so = __dp2a_lo(c0[j], o0, __dp2a_lo(c1[j], o1, __dp2a_lo(c2[j], o2, __dp2a_lo(c3[j], o3, 32768u))));
                ev |= (se >> 16) << (8 * j);
                od |= (so >> 16) << (8 * j);
            }
            const int y = Y0 + 2 * m;
            if (y < L.h) *reinterpret_cast<uint32_t*>(dst + (size_t)y * L.pitch + x) = ev;
            if (y + 1 < L.h) *reinterpret_cast<uint32_t*>(dst + (size_t)(y + 1) * L.pitch + x) = od;
            p0 = p1; p1 = p2; p2 = p3;
        }
    }
}

int launch_blur(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int n, cudaStream_t st) {
    blur_kernel<<<dim3(hg.ntiles, n), 256, 0, st>>>(db.geom, db.tiles, fs, db.pyr, db.blur);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
