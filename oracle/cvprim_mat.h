// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Restatement of the float cv::Mat arithmetic that the reference's matcher / Frame / KeyFrame / MapPoint code
// reaches through OpenCV's matrix expressions (OpenCV is external and un-pinned, /root/reference/CMakeLists.txt:31-37).
// Pinned against Python cv2 4.13.0 by tests/test_oracle_vs_cv2.py (gemm incl. transposed / scaled forms, norm,
// dot, convertTo scaling, scaleAdd, undistortPoints).
//
// Call sites in the reference (examples): Rcw*x3Dw+tcw src/ORBmatcher.cc:326,855,1014,1363,1500; -Rcw.t()*tcw 305,992,
// 1343,1480; sRcw/scw 303-304; cv::norm(PO) 348,877,1036,1516, src/Frame.cc:300; PO.dot(Pn) src/Frame.cc:308,
// src/ORBmatcher.cc:356; cv::undistortPoints src/Frame.cc:422,448; cv::norm(IL,IR,NORM_L1) src/Frame.cc:584.
#pragma once
#include <cmath>
#include <cstddef>

namespace cvprim {

enum { GEMM_A_T = 1, GEMM_B_T = 2 };

// D = alpha*op(A)*op(B) + beta*C for CV_32F. A is ar x ac (before op), B is br x bc; steps in floats. C may be null.
// OpenCV semantics (modules/core/src/matmul.simd.hpp, 4.x): without transposes and with an inner length of 2..4 that
// equals one of the output dimensions, the dot products are accumulated left to right in FLOAT and combined as
// (float)(t*alpha + c*beta) in double; every other shape takes the generic kernel: products and sum in DOUBLE, in
// order, then (float)(s*alpha + c*beta).
static inline void gemm32f(const float* A, int ar, int ac, size_t as, const float* B, int br, int bc, size_t bs, double alpha,
                           const float* C, size_t cs, double beta, float* D, size_t ds, int flags) {
    const int m = (flags & GEMM_A_T) ? ac : ar, len = (flags & GEMM_A_T) ? ar : ac;
    const int n = (flags & GEMM_B_T) ? br : bc;
    (void)br;
    const bool small = flags == 0 && len >= 2 && len <= 4 && (len == n || len == m);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double r;
            if (small) {
                float t = A[i * as + 0] * B[0 * bs + j];
                for (int k = 1; k < len; ++k) t = t + A[i * as + k] * B[k * bs + j];
                r = (double)t * alpha;
            } else {
                double s = 0;
                for (int k = 0; k < len; ++k) {
                    const float a = (flags & GEMM_A_T) ? A[k * as + i] : A[i * as + k];
                    const float b = (flags & GEMM_B_T) ? B[j * bs + k] : B[k * bs + j];
                    s += (double)a * (double)b;
                }
                r = s * alpha;
            }
            if (C) r += (double)C[i * cs + j] * beta;
            D[i * ds + j] = (float)r;
        }
}

// cv::norm(m) (NORM_L2) of a CV_32F matrix: squares accumulated in double in storage order, double sqrt.
static inline double norm_l2_32f(const float* p, int rows, int cols, size_t step) {
    double s = 0;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) { const double v = p[y * step + x]; s += v * v; }
    return std::sqrt(s);
}
// cv::norm(a, b, NORM_L1) of CV_32F matrices: sum of |a-b| (float difference) in double.
static inline double norm_l1_diff_32f(const float* a, size_t as, const float* b, size_t bs, int rows, int cols) {
    double s = 0;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) s += (double)std::fabs(a[y * as + x] - b[y * bs + x]);
    return s;
}
// cv::Mat::dot for CV_32F: products and sum in double, storage order.
static inline double dot_32f(const float* a, size_t as, const float* b, size_t bs, int rows, int cols) {
    double s = 0;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) s += (double)a[y * as + x] * (double)b[y * bs + x];
    return s;
}

// cv::undistortPoints(src, dst, K, dist, noArray(), P=K) for CV_32FC2 points, K CV_32F 3x3, dist = (k1,k2,p1,p2[,k3])
// CV_32F: OpenCV 4.x cvUndistortPointsInternal with the default criteria (5 iterations, no epsilon test inside the
// first 5), all arithmetic in double, result stored as float.
static inline void undistort_points_32f(const float* src, float* dst, int n, const float* K, size_t ks, const float* dist, int nd) {
    double k[14] = {0};
    for (int i = 0; i < nd && i < 14; ++i) k[i] = dist[i];
    const double fx = K[0], fy = K[ks + 1], cx = K[2], cy = K[ks + 2];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; ++i) {
        double x = src[2 * i], y = src[2 * i + 1];
        const double u = x, v = y;
        (void)u; (void)v;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        // P = K: xx = fx*x + cx (the 3x3 product with z = 1, row by row in double)
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1);
        dst[2 * i] = (float)(xx * ww);
        dst[2 * i + 1] = (float)(yy * ww);
    }
}

}  // namespace cvprim
