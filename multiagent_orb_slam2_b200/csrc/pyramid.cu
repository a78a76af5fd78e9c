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
// resize: one block (4 warps) = one 128 x 32 output tile, one warp = 8 consecutive output rows of it, one
// lane = 4 output columns (measured at B=512: 0.381 / 0.365 / 0.362 / 0.364 ms with 64 / 32 / 16 / 8 tile rows). The source window of the tile arrives in shared memory by one TMA box load
// (rs_bw x rs_bh bytes, zero fill beyond the level - never read); while it is in flight each thread
// fetches the column taps of its 4 columns and the row taps of its 8 rows into registers.
// The horizontal pass of a source row is kept for the next output row: consecutive output rows share
// a source row (the second tap of row y is the first tap of row y+1 unless the scale skips a row), so
// a warp interpolates ~1.2 source rows per output row instead of 2. Row decisions are warp-uniform
// (except in the narrow tiles of a level's last column, see below).
constexpr int kRsTW = 128, kRsTH = 32, kRsRowsPerWarp = 8, kRsThreads = 32 * kRsTH / kRsRowsPerWarp;

// PRMT with a run-time selector whose nibbles are all < 8 (__byte_perm masks the selector with 0x7777 first: one LOP3 per use)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__global__ void __launch_bounds__(kRsThreads)
resize_kernel(const Geometry* __restrict__ g, const LinTap* __restrict__ taps, const __grid_constant__ TmaMaps maps,
              uint8_t* __restrict__ pyr, int level) {
    extern __shared__ __align__(128) uint8_t win[];  // rs_bh rows of rs_bw bytes: the TMA box
    __shared__ __align__(8) uint64_t bar;
    const LevelGeom& L = g->lv[level];
    const LevelGeom& S = g->lv[level - 1];
    const int frame = blockIdx.z;
    const int X0 = blockIdx.x * kRsTW, Y0 = blockIdx.y * kRsTH;
    uint8_t* dst = pyr + (size_t)frame * g->pyr_bytes + L.img_off;
    const LinTap* tx = taps + L.tab_x_off;
    const LinTap* ty = taps + L.tab_y_off;
    const int bw = g->rs_bw;

    // source window origin (taps are monotone); the box covers the whole tile at any supported scale
    const int sx_lo = tx[X0].ofs & ~15, sy_lo = ty[Y0].ofs;  // innermost TMA coordinate: 16-byte granular
    // Programmatic dependent launch (launch_resize_level(..., pdl)): the next level's grid may be scheduled
    // while this one runs - its blocks get as far as the wait below (set-up, table reads) - and this grid's own source level
    // is only touched after the previous grid has completed. Both instructions are no-ops in an ordinary launch.
    pdl_release_dependents();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        pdl_wait();
        mbar_expect_tx(&bar, (uint32_t)(bw * g->rs_bh));
        tma_load_3d(win, &maps.m[level - 1], sx_lo, sy_lo, frame, &bar);
    }
    // all table reads of this thread, issued while the window is in flight
    // A tile of the level's last column may hold only a few column quads: its warps then put 2 or 4 row groups
    // side by side (16 or 8 lanes per row, each lane group a contiguous run of 4 or 2 rows) instead of running
    // 8 rows with most lanes idle.
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int nqv = min(32, (L.w - X0 + 3) >> 2);
    const int lpr_log = nqv <= 8 ? 3 : nqv <= 16 ? 4 : 5;         // lanes per row: 8, 16 or 32
    const int rpg = kRsRowsPerWarp >> (5 - lpr_log);                // rows per lane group: 2, 4 or 8
    const int q = lane & ((1 << lpr_log) - 1);
    const int x0 = X0 + 4 * q, yw = Y0 + kRsRowsPerWarp * rg + (lane >> lpr_log) * rpg;
    // a LinTap is 8 bytes (source index | two 16-bit weights): one 64-bit load each
    uint2 tcol[4], trow[kRsRowsPerWarp];
#pragma unroll
    for (int k = 0; k < 4; ++k) tcol[k] = reinterpret_cast<const uint2*>(tx)[min(x0 + k, L.w - 1)];
#pragma unroll
    for (int rr = 0; rr < kRsRowsPerWarp; ++rr) trow[rr] = reinterpret_cast<const uint2*>(ty)[min(yw + rr, L.h - 1)];
    __syncthreads();
    mbar_wait(&bar, 0);

    if (x0 >= L.w) return;
    // Per thread constants of its 4 columns: the source bytes (sx, sx+1) of all four lie within 8 bytes
    // of column 0's first tap (scale < 2), i.e. in 3 words w0 w1 w2 of a window row. Two funnel shifts
    // bring those 8 bytes to (u0, u1); for column k one byte permute picks its two taps and one IDP.2A
    // applies the two 11-bit weights (u16 pair) to the two pixels (u8 pair).
    const int o0 = (int)tcol[0].x - sx_lo;
    uint32_t wofs = (uint32_t)(o0 & ~3), shbits = (uint32_t)(o0 & 3) * 8u;
    uint32_t sel[4], coef[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t d = tcol[k].x - tcol[0].x;  // 0..6
        sel[k] = d | (d + 1) << 4 | 0x4400u;       // only the two low bytes of the permute result are used by IDP.2A (lo variant)
        coef[k] = tcol[k].y;                        // c0 | c1 << 16
        asm volatile("" : "+r"(sel[k]));            // keep in a register (ptxas re-derived it in every row otherwise)
    }
    asm volatile("" : "+r"(wofs), "+r"(shbits));
    // horizontal pass of window row s for the 4 columns, already shifted right by 4 as the vertical pass wants it
    auto horiz = [&](int s, uint32_t (&h)[4]) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(win + (uint32_t)(s - sy_lo) * (uint32_t)bw + wofs);
        const uint32_t a0 = r[0], a1 = r[1], a2 = r[2];
        const uint32_t u0 = __funnelshift_r(a0, a1, shbits), u1 = __funnelshift_r(a1, a2, shbits);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __dp2a_lo(coef[k], prmt(u0, u1, sel[k]), 0u) >> 4;
    };
    int crow = -1;
    uint32_t hc[4] = {0, 0, 0, 0};  // cached horizontal pass: source row and values
    uint8_t* drow = dst + (size_t)yw * L.pitch + x0;
    const int dpitch = L.pitch, sh_max = S.h - 1, nrows = min(rpg, L.h - yw);
#pragma unroll
    for (int rr = 0; rr < kRsRowsPerWarp; ++rr) {
        if (rr >= nrows) break;
        const int s0 = (int)trow[rr].x, s1 = min(s0 + 1, sh_max);
        uint32_t h0[4], h1[4];
        if (s0 == crow) {
#pragma unroll
            for (int k = 0; k < 4; ++k) h0[k] = hc[k];
        } else {
            horiz(s0, h0);
        }
        if (s1 == s0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) h1[k] = h0[k];
        } else {
            horiz(s1, h1);
        }
        crow = s1;
        // (b * h) >> 16 as the high word of (b << 16) * h: weights are 0..2048, h < 2^15
        const uint32_t b0 = trow[rr].y << 16, b1 = trow[rr].y & 0xffff0000u;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hc[k] = h1[k];
            v[k] = (__umulhi(b0, h0[k]) + __umulhi(b1, h1[k]) + 2u) >> 2;  // <= 255: the weights of each pass sum to 2048
        }
        // four results, each in the low byte of its register -> one word: 3 byte permutes
        const uint32_t packed = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
        // rows are 128-byte pitched and x0 is a multiple of 4: aligned 32-bit store (pitch slack absorbs the tail)
        *reinterpret_cast<uint32_t*>(drow) = packed;
        drow += dpitch;
    }
}

int launch_resize_level(const Geometry& hg, const DeviceBuffers& db, const TmaMaps& maps, int level, int n, cudaStream_t st, bool pdl) {
    const LevelGeom& L = hg.lv[level];
    dim3 grid(ceil_div(L.w, kRsTW), ceil_div(L.h, kRsTH), n);
    if (pdl)  // overlap this level's launch and set-up with the tail of the level before it (same stream)
        ORB_CUDA_TRY(launch_pdl(resize_kernel, grid, dim3(kRsThreads), (size_t)hg.rs_bw * hg.rs_bh, st, db.geom, db.taps, maps, db.pyr, level));
    else
        resize_kernel<<<grid, kRsThreads, (size_t)hg.rs_bw * hg.rs_bh, st>>>(db.geom, db.taps, maps, db.pyr, level);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// ------------------------------------------------------------------------------------------------
// blur: separable {18,34,48,56,48,34,18}, single rounding (sum + 2^15) >> 16, BORDER_REFLECT_101.
// One block = one 128 x 58 output tile of one level of one frame (input 136 x 64 incl. halo).
//   load : one TMA box (144 x 64 bytes) into shared memory; border tiles patch the reflected bytes
//   h    : 4 outputs per thread from 3 words: the taps laid over the words at each output's byte offset, 10 x IDP.4A per
//          4 outputs (u8 x u8 -> u32, exact); rows are processed in vertical pairs and stored as
//          (row 2m | row 2m+1 << 16) so that
//   v    : one IDP.2A covers two vertical taps: 4 x IDP.2A per output, accumulator preloaded with
//          the rounding constant; 4 columns per thread, sliding window down the tile.
// ~14 instructions per pixel instead of ~120 for the straightforward byte-wise version.
constexpr int kBlurTW = 128, kBlurTH = 58, kBlurInRows = 64, kBlurInWords = 40;  // 160-byte TMA box rows
constexpr int kBlurHsvPitch = kBlurTW + 4;  // +4 words: rows of a partial tile's flattened items fall into different banks
constexpr int kBlurLead = 16;  // bytes left of the tile in the box: innermost TMA coordinate is 16-byte granular (3 needed)

__device__ __forceinline__ int reflect101(int p, int n) {
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * n - 2 - p : p;
    return max(0, min(p, n - 1));  // only out-of-tile lanes of partial tiles ever hit the clamp
}

__global__ void __launch_bounds__(256)
blur_kernel(const Geometry* __restrict__ g, const BlurTile* __restrict__ tiles, const __grid_constant__ TmaMaps maps,
            uint8_t* __restrict__ blur) {
    __shared__ __align__(128) uint32_t in_w[kBlurInRows][kBlurInWords];
    __shared__ __align__(16) uint32_t hsv[kBlurInRows / 2][kBlurHsvPitch];
    __shared__ __align__(8) uint64_t bar;
    const BlurTile t = tiles[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelGeom& L = g->lv[t.level];
    uint8_t* dst = blur + (size_t)frame * g->blur_bytes + L.blur_off;
    const int X0 = t.tx * kBlurTW, Y0 = t.ty * kBlurTH;
    const int tid = threadIdx.x;

    // input tile: columns X0-16 .. X0+143, rows Y0-3 .. Y0+60, zero filled outside the level by the TMA unit
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, (uint32_t)(kBlurInRows * kBlurInWords * 4));
        tma_load_3d(in_w, &maps.m[t.level], X0 - kBlurLead, Y0 - 3, frame, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    // BORDER_REFLECT_101: tiles that touch the level's frame rebuild their out-of-level bytes from the
    // in-level ones (which are inside the same tile: the reflection reaches at most 4 px / 3 rows in)
    {
        constexpr int BW = kBlurInWords * 4;
        uint8_t* inb = reinterpret_cast<uint8_t*>(&in_w[0][0]);
        const int x_lo = X0 - kBlurLead, y_lo = Y0 - 3;
        const bool top = Y0 == 0, left = X0 == 0, bottom = y_lo + kBlurInRows > L.h, right = x_lo + BW > L.w;
        if (top || left || bottom || right) {
            // only the 3 rows / 3-4 columns next to the frame can reach an output of this tile
            auto fix = [&](int r, int c) {
                const int y = y_lo + r, x = x_lo + c;
                const int sr = reflect101(y, L.h) - y_lo, sc = reflect101(x, L.w) - x_lo;
                if (r < kBlurInRows && c < BW && sr >= 0 && sr < kBlurInRows && sc >= 0 && sc < BW) inb[r * BW + c] = inb[sr * BW + sc];
            };
            if (top) for (int i = tid; i < 3 * BW; i += 256) fix(i / BW, i % BW);
            if (bottom) for (int i = tid; i < 3 * BW; i += 256) fix(L.h - y_lo + i / BW, i % BW);
            if (left) for (int i = tid; i < 4 * kBlurInRows; i += 256) fix(i >> 2, kBlurLead - 4 + (i & 3));
            if (right) for (int i = tid; i < 4 * kBlurInRows; i += 256) fix(i >> 2, L.w - x_lo + (i & 3));
            __syncthreads();
        }
    }

    // Partial tiles (right column / bottom row of a level) only enumerate what lies inside the level: nq column
    // quads and nrp input row pairs; items are flattened so that whole warps drop out instead of lanes.
    const int nq = min(kBlurTW / 4, (L.w - X0 + 3) >> 2);
    const int nrp = min(kBlurInRows / 2, (min(kBlurTH, L.h - Y0) + 6 + 1) >> 1);  // output rows + 6 halo rows, in pairs
    const uint32_t qinv = (65536u + nq - 1) / nq;  // floor(i / nq) == i * qinv >> 16 for i * nq < 65536
    // horizontal pass: item = (row pair, column quad)
    for (int it = tid; it < nrp * nq; it += 256) {
        const int pr = (int)((uint32_t)it * qinv >> 16), k = it - pr * nq;
        uint32_t h[2][4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const uint32_t* wr = &in_w[2 * pr + rr][kBlurLead / 4 - 1 + k];  // word of bytes X0+4k-4 .. X0+4k-1
            const uint32_t w0 = wr[0], w1 = wr[1], w2 = wr[2];
            // output j (column 4k+j) uses bytes j+1 .. j+7 of the 12-byte string w0 w1 w2. The data words stay where they
            // are and the seven taps are laid over them at the byte offset of each output: 10 IDP.4A per 4 outputs instead of
            // 6 funnel shifts + 8 IDP.4A (the shifts ran on the ALU pipe, the busier one of this kernel)
            constexpr uint32_t T18 = 18u, T34 = 34u, T48 = 48u, T56 = 56u;
            h[rr][0] = __dp4a(w0, T18 << 8 | T34 << 16 | T48 << 24, __dp4a(w1, T56 | T48 << 8 | T34 << 16 | T18 << 24, 0u));
            h[rr][1] = __dp4a(w0, T18 << 16 | T34 << 24, __dp4a(w1, T48 | T56 << 8 | T48 << 16 | T34 << 24, __dp4a(w2, T18, 0u)));
            h[rr][2] = __dp4a(w0, T18 << 24, __dp4a(w1, T34 | T48 << 8 | T56 << 16 | T48 << 24, __dp4a(w2, T34 | T18 << 8, 0u)));
            h[rr][3] = __dp4a(w1, T18 | T34 << 8 | T48 << 16 | T56 << 24, __dp4a(w2, T48 | T34 << 8 | T18 << 16, 0u));
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
    const int seg = (int)((uint32_t)tid * qinv >> 16), quad = tid - seg * nq;  // (tid / nq, tid % nq): 8 segments of 4 row pairs
    const int x = X0 + 4 * quad;
    if (seg < 8 && Y0 + 8 * seg < L.h) {
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
            uint32_t se[4], so[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                se[j] = __dp2a_lo(c0[j], e0, __dp2a_lo(c1[j], e1, __dp2a_lo(c2[j], e2, __dp2a_lo(c3[j], e3, 32768u))));
                so[j] = __dp2a_lo(c0[j], o0, __dp2a_lo(c1[j], o1, __dp2a_lo(c2[j], o2, __dp2a_lo(c3[j], o3, 32768u))));
            }
            // the result is (sum + 2^15) >> 16 <= 255, i.e. byte 2 of the accumulator: four of them -> one word by 3 byte permutes
            const uint32_t ev = __byte_perm(__byte_perm(se[0], se[1], 0x0062), __byte_perm(se[2], se[3], 0x0062), 0x5410);
            const uint32_t od = __byte_perm(__byte_perm(so[0], so[1], 0x0062), __byte_perm(so[2], so[3], 0x0062), 0x5410);
            const int y = Y0 + 2 * m;
            if (y < L.h) *reinterpret_cast<uint32_t*>(dst + (size_t)y * L.pitch + x) = ev;
            if (y + 1 < L.h) *reinterpret_cast<uint32_t*>(dst + (size_t)(y + 1) * L.pitch + x) = od;
            p0 = p1; p1 = p2; p2 = p3;
        }
    }
}

int launch_blur(const Geometry& hg, const DeviceBuffers& db, const TmaMaps& maps, int n, cudaStream_t st) {
    blur_kernel<<<dim3(hg.ntiles, n), 256, 0, st>>>(db.geom, db.tiles, maps, db.blur);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
