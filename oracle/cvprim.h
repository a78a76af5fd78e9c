// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Restatement of the OpenCV primitives that the reference's ORB frontend calls but does not
// vendor (OpenCV is an external, un-pinned dependency: /root/reference/CMakeLists.txt:31-37).
// Pinned against Python cv2 4.13.0 by tests/test_oracle_vs_cv2.py (bit-exact).
//
// Call sites in the reference that these replace:
//   cv::resize(INTER_LINEAR, 8UC1)        src/ORBextractor.cc:1120
//   cv::copyMakeBorder(REFLECT_101)       src/ORBextractor.cc:1122-1128
//   cv::GaussianBlur(7x7, sigma 2)        src/ORBextractor.cc:1086
//   cv::FAST(img, kps, th, nonmax=true)   src/ORBextractor.cc:809-815
//   cv::fastAtan2                         src/ORBextractor.cc:103
//   cvRound / cvFloor / cvCeil            src/ORBextractor.cc:81,115,119-120,442,456-460,1112
//
// All functions work on raw (pointer, stride) views so that both the plain oracle
// (orb_oracle.cc) and the cv:: compat shim (oracle/shim) can share them.
#pragma once
#include <cfloat>
#include <cstddef>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace cvprim {

// round-half-to-even, as cvRound (lrint under the default rounding mode)
static inline int round_half_even(double v) { return (int)std::nearbyint(v); }
static inline int round_half_even(float v) { return (int)std::nearbyintf(v); }
static inline int ifloor(double v) { int i = (int)v; return i - (i > v); }
static inline int iceil(double v) { int i = (int)v; return i + (i < v); }

// BORDER_REFLECT_101 index: -1 -> 1, n -> n-2 (no edge repeat)
static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * n - 2 - p;
    }
    return p;
}

// ---------------------------------------------------------------------------------------
// resize, INTER_LINEAR, 8-bit single channel. 11-bit fixed point coefficients.
struct LinearTab {
    std::vector<int> ofs;        // source index of the left/top tap
    std::vector<short> c0, c1;   // 2048-scaled weights
};

static inline LinearTab linear_table(int src_n, int dst_n) {
    LinearTab t;
    t.ofs.resize(dst_n); t.c0.resize(dst_n); t.c1.resize(dst_n);
    const double scale = (double)src_n / dst_n;
    for (int d = 0; d < dst_n; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = ifloor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
        t.ofs[d] = s;
        t.c0[d] = (short)round_half_even((1.f - f) * 2048.f);
        t.c1[d] = (short)round_half_even(f * 2048.f);
    }
    return t;
}

static inline void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                                    uint8_t* dst, int dw, int dh, size_t dstride) {
    LinearTab tx = linear_table(sw, dw), ty = linear_table(sh, dh);
    std::vector<int> r0(dw), r1(dw);
    for (int y = 0; y < dh; ++y) {
        const int sy0 = ty.ofs[y];
        const int sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        const uint8_t* p0 = src + (size_t)sy0 * sstride;
        const uint8_t* p1 = src + (size_t)sy1 * sstride;
        for (int x = 0; x < dw; ++x) {
            const int sx0 = tx.ofs[x];
            const int sx1 = sx0 + 1 < sw ? sx0 + 1 : sw - 1;
            r0[x] = p0[sx0] * tx.c0[x] + p0[sx1] * tx.c1[x];
            r1[x] = p1[sx0] * tx.c0[x] + p1[sx1] * tx.c1[x];
        }
        const int b0 = ty.c0[y], b1 = ty.c1[y];
        uint8_t* o = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; ++x)
            o[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------------------------------
// GaussianBlur 7x7 sigma=2 on 8U: fixed-point kernel {18,34,48,56,48,34,18}/256 applied
// separably with a single rounding: (sum + 2^15) >> 16. Border REFLECT_101. In-place safe.
static const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};

static inline void gaussian7x7_u8(const uint8_t* src, int w, int h, size_t sstride,
                                  uint8_t* dst, size_t dstride) {
    // horizontal pass into 16-bit rows (sum <= 255*256 fits), reflected indices only in the 3-pixel margins;
    // vertical pass accumulates whole rows (tap-major), which the compiler vectorises
    static thread_local std::vector<uint16_t> hbuf;
    static thread_local std::vector<uint32_t> acc;
    hbuf.resize((size_t)w * h);
    acc.resize(w);
    uint16_t* const hb = hbuf.data();   // thread_local: hoisted out of the loops
    uint32_t* const ac = acc.data();
    for (int y = 0; y < h; ++y) {
        const uint8_t* p = src + (size_t)y * sstride;
        uint16_t* o = hb + (size_t)y * w;
        const int x_lo = w < 7 ? w : 3, x_hi = w < 7 ? w : w - 3;
        for (int x = 0; x < x_lo; ++x) {
            int s = 0;
            for (int k = 0; k < 7; ++k) s += kGauss7[k] * p[reflect101(x + k - 3, w)];
            o[x] = (uint16_t)s;
        }
        for (int x = x_lo; x < x_hi; ++x)
            o[x] = (uint16_t)(18 * (p[x - 3] + p[x + 3]) + 34 * (p[x - 2] + p[x + 2]) + 48 * (p[x - 1] + p[x + 1]) + 56 * p[x]);
        for (int x = x_hi > x_lo ? x_hi : x_lo; x < w; ++x) {
            int s = 0;
            for (int k = 0; k < 7; ++k) s += kGauss7[k] * p[reflect101(x + k - 3, w)];
            o[x] = (uint16_t)s;
        }
    }
    for (int y = 0; y < h; ++y) {
        uint32_t* a = ac;
        for (int x = 0; x < w; ++x) a[x] = 32768u;
        for (int k = 0; k < 7; ++k) {
            const uint16_t* r = hb + (size_t)reflect101(y + k - 3, h) * w;
            const uint32_t c = (uint32_t)kGauss7[k];
            for (int x = 0; x < w; ++x) a[x] += c * r[x];
        }
        uint8_t* o = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) o[x] = (uint8_t)(a[x] >> 16);
    }
}

// ---------------------------------------------------------------------------------------
// copyMakeBorder with BORDER_REFLECT_101 (isolated: reflects about the source ROI edge).
static inline void copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstride,
                                               uint8_t* dst, size_t dstride, int top, int bottom,
                                               int left, int right) {
    const int dw = w + left + right, dh = h + top + bottom;
    std::vector<uint8_t> row(dw);
    // rows are produced into a scratch first so that src may alias the interior of dst
    std::vector<uint8_t> out((size_t)dw * dh);
    for (int y = 0; y < dh; ++y) {
        const uint8_t* p = src + (size_t)reflect101(y - top, h) * sstride;
        uint8_t* o = &out[(size_t)y * dw];
        for (int x = 0; x < dw; ++x) o[x] = p[reflect101(x - left, w)];
    }
    for (int y = 0; y < dh; ++y) std::memcpy(dst + (size_t)y * dstride, &out[(size_t)y * dw], dw);
}

// ---------------------------------------------------------------------------------------
// FAST-9/16. Ring offsets in OpenCV's order (dx,dy).
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Largest threshold t (minus nothing) for which the pixel is a FAST-9 corner is score:
// corner at threshold th  <=>  fast_score >= th. Returns -1 when no arc has a positive margin.
static inline int fast_score(const uint8_t* p, size_t stride) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[(std::ptrdiff_t)kRingDy[k] * (std::ptrdiff_t)stride + kRingDx[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    // the arcs starting at k and k+1 share the 8 values d[k+1..k+8]: one pass over the even k covers all 16 arcs
    int best = 0;
    for (int k = 0; k < 16; k += 2) {
        int mn = d[k + 1], mx = d[k + 1];
        for (int j = 2; j <= 8; ++j) {
            const int e = d[k + j];
            mn = e < mn ? e : mn;
            mx = e > mx ? e : mx;
        }
        // all brighter-than-ring by mn, or all darker-than-ring by -mx
        const int a = d[k], b = d[k + 9];
        const int mn0 = a < mn ? a : mn, mn1 = b < mn ? b : mn;
        const int mx0 = a > mx ? a : mx, mx1 = b > mx ? b : mx;
        if (mn0 > best) best = mn0;
        if (mn1 > best) best = mn1;
        if (-mx0 > best) best = -mx0;
        if (-mx1 > best) best = -mx1;
    }
    return best - 1;
}

struct FastKp { int x, y, score; };

// cv::FAST(img, kps, threshold, true) on a w x h image: corners with score >= threshold in the
// interior [3,w-3)x[3,h-3), kept iff strictly greater than all 8 neighbours of the thresholded
// score map (0 outside the interior / for non-corners). Output row-major.
// Like OpenCV's FAST_t<16> the pixel is first put through the cheap necessary condition - an arc of 9 contiguous ring
// pixels contains one pixel of every opposite pair (k, k+8), so for a corner at threshold t every pair must hold a pixel
// brighter than v+t, or every pair one darker than v-t; pairs (0,8), (4,12), (2,10), (6,14) are tested - and only the
// survivors (a few per cent) get the exact 16-arc score. The result is identical to scoring every pixel.
static inline void fast9_nms(const uint8_t* img, int w, int h, size_t stride, int threshold,
                             std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    static thread_local std::vector<int> sc;
    sc.assign((size_t)w * h, 0);
    int* const scp = sc.data();   // thread_local objects are reached through a call in a shared library: hoist
    const std::ptrdiff_t st = (std::ptrdiff_t)stride;
    const std::ptrdiff_t o0 = 3 * st, o8 = -3 * st, o4 = 3, o12 = -3, o2 = 2 * st + 2, o10 = -2 * st - 2, o6 = -2 * st + 2, o14 = 2 * st - 2;
    static thread_local std::vector<uint8_t> weak;
    weak.resize(w);
    uint8_t* const wk = weak.data();
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* row = img + (size_t)y * stride;
        int* srow = scp + (size_t)y * w;
        // row-wide weak form of the first two pair tests (|v - p| > t for one pixel of the pair, either sign): a
        // superset of the survivors, written so that the compiler vectorises it
        {
            const uint8_t *r0 = row + o0, *r8 = row + o8;
            uint8_t* m = wk;
            for (int x = 3; x < w - 3; ++x) {
                const int v = row[x];
                const int d0 = v - r0[x], d8 = v - r8[x], d4 = v - row[x + 3], d12 = v - row[x - 3];
                const int a0 = d0 < 0 ? -d0 : d0, a8 = d8 < 0 ? -d8 : d8, a4 = d4 < 0 ? -d4 : d4, a12 = d12 < 0 ? -d12 : d12;
                m[x] = (uint8_t)(((a0 > threshold) | (a8 > threshold)) & ((a4 > threshold) | (a12 > threshold)));
            }
        }
        for (int x = 3; x < w - 3; ++x) {
            if (!wk[x]) continue;
            const uint8_t* p = row + x;
            const int hi = p[0] + threshold, lo = p[0] - threshold;
            const int a = p[o0], b = p[o8];
            int bright = (a > hi) | (b > hi), dark = (a < lo) | (b < lo);
            const int c = p[o4], d = p[o12];
            bright &= (c > hi) | (d > hi); dark &= (c < lo) | (d < lo);
            if (!(bright | dark)) continue;
            const int e = p[o2], f = p[o10];
            bright &= (e > hi) | (f > hi); dark &= (e < lo) | (f < lo);
            if (!(bright | dark)) continue;
            const int g = p[o6], i = p[o14];
            bright &= (g > hi) | (i > hi); dark &= (g < lo) | (i < lo);
            if (!(bright | dark)) continue;
            const int s = fast_score(p, stride);
            if (s >= threshold) srow[x] = s;
        }
    }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = scp[(size_t)y * w + x];
            if (s <= 0) continue;
            bool mx = true;
            for (int dy = -1; dy <= 1 && mx; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    if (scp[(size_t)(y + dy) * w + x + dx] >= s) { mx = false; break; }
                }
            if (mx) out.push_back({x, y, s});
        }
}

// ---------------------------------------------------------------------------------------
// cv::fastAtan2 (degrees, fp32, no fused multiply-add; compile with -ffp-contract=off)
static inline float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

}  // namespace cvprim
