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
    std::vector<uint16_t> hbuf((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* p = src + (size_t)y * sstride;
        for (int x = 0; x < w; ++x) {
            int s = 0;
            for (int k = 0; k < 7; ++k) s += kGauss7[k] * p[reflect101(x + k - 3, w)];
            hbuf[(size_t)y * w + x] = (uint16_t)s;
        }
    }
    for (int y = 0; y < h; ++y) {
        uint8_t* o = dst + (size_t)y * dstride;
        const uint16_t* rows[7];
        for (int k = 0; k < 7; ++k) rows[k] = &hbuf[(size_t)reflect101(y + k - 3, h) * w];
        for (int x = 0; x < w; ++x) {
            uint32_t s = 32768u;
            for (int k = 0; k < 7; ++k) s += (uint32_t)kGauss7[k] * rows[k][x];
            o[x] = (uint8_t)(s >> 16);
        }
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
    int d[16];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[(std::ptrdiff_t)kRingDy[k] * (std::ptrdiff_t)stride + kRingDx[k]];
    int best = 0;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) {
            const int e = d[(k + j) & 15];
            mn = e < mn ? e : mn;
            mx = e > mx ? e : mx;
        }
        // all brighter-than-ring by mn, or all darker-than-ring by -mx
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best - 1;
}

struct FastKp { int x, y, score; };

// cv::FAST(img, kps, threshold, true) on a w x h image: corners with score >= threshold in the
// interior [3,w-3)x[3,h-3), kept iff strictly greater than all 8 neighbours of the thresholded
// score map (0 outside the interior / for non-corners). Output row-major.
static inline void fast9_nms(const uint8_t* img, int w, int h, size_t stride, int threshold,
                             std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = fast_score(img + (size_t)y * stride + x, stride);
            sc[(size_t)y * w + x] = s >= threshold ? s : 0;
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = sc[(size_t)y * w + x];
            if (s <= 0) continue;
            bool mx = true;
            for (int dy = -1; dy <= 1 && mx; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    if (sc[(size_t)(y + dy) * w + x + dx] >= s) { mx = false; break; }
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
