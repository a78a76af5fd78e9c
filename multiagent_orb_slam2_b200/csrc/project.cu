// Map-point projection for the projection-guided searches: Frame::isInFrustum
// (/root/reference/src/Frame.cc:269-325) and the projection prologue shared by ORBmatcher::Fuse / SearchByProjection(KF, Scw) /
// SearchBySim3 (src/ORBmatcher.cc:849-889, 323-363) for a whole batch of map points held as structure-of-arrays in HBM.
// One thread per map point; outputs feed orbm_window_knn2_device / orbm_window_lists_device directly (window centre,
// radius, level window), so the projection -> window -> Hamming chain needs no host round trip.
//
// Arithmetic contract (what the reference's cv::Mat expressions evaluate to with OpenCV 4.13, pinned in
// tests/test_projection_oracle.py against cv2.gemm / cv2.norm):
//   Pc   = Rcw*P + tcw     float, each row ((r0*p0 + r1*p1) + r2*p2) + t, no fused multiply-add
//   dist = cv::norm(P-Ow)  double accumulation of the squares in order, double sqrt, rounded to float
//   PO.dot(Pn)             double accumulation of the float products in order
//   PredictScale           ceil(logf(mfMaxDistance/dist) / mfLogScaleFactor) in float (std::log(float) is the overload in
//                          scope: Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:36 has `using namespace std`); the device takes
//                          the correctly rounded float of a double log.
#include "common.cuh"

namespace orb {

struct ProjectParams {
    orbm_camera cam;
    float cos_limit;
    float th;
    int mode;
    int n;
};

__global__ void __launch_bounds__(256) project_points_kernel(const ProjectParams p, const float* __restrict__ pos,
                                                             const float* __restrict__ normal, const float* __restrict__ max_dist,
                                                             const float* __restrict__ min_dist, uint8_t* __restrict__ alive,
                                                             float* __restrict__ out_u, float* __restrict__ out_v,
                                                             float* __restrict__ out_ur, int32_t* __restrict__ out_level,
                                                             float* __restrict__ out_cos, float* __restrict__ out_radius,
                                                             int32_t* __restrict__ out_min_level, int32_t* __restrict__ out_max_level) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const orbm_camera& c = p.cam;
    const float P0 = pos[3 * i], P1 = pos[3 * i + 1], P2 = pos[3 * i + 2];
    const float x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.Rcw[0], P0), __fmul_rn(c.Rcw[1], P1)), __fmul_rn(c.Rcw[2], P2)), c.tcw[0]);
    const float y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.Rcw[3], P0), __fmul_rn(c.Rcw[4], P1)), __fmul_rn(c.Rcw[5], P2)), c.tcw[1]);
    const float z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.Rcw[6], P0), __fmul_rn(c.Rcw[7], P1)), __fmul_rn(c.Rcw[8], P2)), c.tcw[2]);
    bool ok = !(z < 0.0f);
    const float invz = __fdiv_rn(1.0f, z);
    float u, v;
    if (p.mode == 0) {  // Frame.cc:287-289: fx*PcX*invz+cx
        u = __fadd_rn(__fmul_rn(__fmul_rn(c.fx, x), invz), c.cx);
        v = __fadd_rn(__fmul_rn(__fmul_rn(c.fy, y), invz), c.cy);
        ok = ok && !(u < c.min_x || u > c.max_x) && !(v < c.min_y || v > c.max_y);
    } else {  // ORBmatcher.cc:858-866: x = Pc.x*invz; u = fx*x+cx; KeyFrame::IsInImage
        u = __fadd_rn(__fmul_rn(c.fx, __fmul_rn(x, invz)), c.cx);
        v = __fadd_rn(__fmul_rn(c.fy, __fmul_rn(y, invz)), c.cy);
        ok = ok && (u >= c.min_x && u < c.max_x && v >= c.min_y && v < c.max_y);
    }
    const float ur = __fsub_rn(u, __fmul_rn(c.bf, invz));
    const float maxd = __fmul_rn(1.2f, max_dist[i]), mind = __fmul_rn(0.8f, min_dist[i]);
    const float o0 = __fsub_rn(P0, c.Ow[0]), o1 = __fsub_rn(P1, c.Ow[1]), o2 = __fsub_rn(P2, c.Ow[2]);
    const double s = __dadd_rn(__dadd_rn(__dmul_rn((double)o0, (double)o0), __dmul_rn((double)o1, (double)o1)), __dmul_rn((double)o2, (double)o2));
    const float dist = (float)sqrt(s);
    ok = ok && !(dist < mind || dist > maxd);
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)o0, (double)normal[3 * i]), __dmul_rn((double)o1, (double)normal[3 * i + 1])),
                                 __dmul_rn((double)o2, (double)normal[3 * i + 2]));
    float view_cos = (float)(dot / (double)dist);
    if (p.mode == 0)
        ok = ok && !(view_cos < p.cos_limit);
    else
        ok = ok && !(dot < 0.5 * (double)dist);
    int level = 0;
    float radius = 0.0f;
    if (ok) {
        const float ratio = __fdiv_rn(max_dist[i], dist);
        const float lg = (float)log((double)ratio);
        level = (int)ceilf(__fdiv_rn(lg, c.log_scale_factor));
        level = level < 0 ? 0 : (level >= c.n_levels ? c.n_levels - 1 : level);
        if (p.mode == 0) {  // RadiusByViewingCos (ORBmatcher.cc:133-139) and the th factor (ORBmatcher.cc:64-67)
            float r = view_cos > 0.998f ? 2.5f : 4.0f;
            if (p.th != 1.0f) r = __fmul_rn(r, p.th);
            radius = __fmul_rn(r, c.scale_factors[level]);
        } else {
            radius = __fmul_rn(p.th, c.scale_factors[level]);
        }
    }
    alive[i] = ok;
    out_u[i] = u; out_v[i] = v; out_ur[i] = ur;
    out_level[i] = level;
    out_cos[i] = view_cos;
    if (out_radius) out_radius[i] = radius;
    if (out_min_level) out_min_level[i] = level - 1;
    if (out_max_level) out_max_level[i] = level;
}

// Frame::ComputeStereoFromRGBD (/root/reference/src/Frame.cc:643-664): the depth image sampled at the (truncated) keypoint
// position gives mvDepth, and mvuRight = x - mbf/depth, for the keypoints an extractor left on the device.
__global__ void __launch_bounds__(256) stereo_from_rgbd_kernel(const orbx_keypoint* __restrict__ kps, const orbx_keypoint* __restrict__ kps_un, const int32_t* __restrict__ count, int cap,
                                                               const float* __restrict__ depth_img, int width, int height, size_t pitch_floats,
                                                               float mbf, float* __restrict__ uright, float* __restrict__ depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    float ur = -1.0f, dz = -1.0f;
    if (i < min(*count, cap)) {
        const orbx_keypoint kp = kps[i];
        const int u = (int)kp.x, v = (int)kp.y;  // imDepth.at<float>(v,u) with float arguments: truncation
        if (u >= 0 && u < width && v >= 0 && v < height) {
            const float d = depth_img[(size_t)v * pitch_floats + u];
            if (d > 0) {
                dz = d;
                ur = __fsub_rn(kps_un ? kps_un[i].x : kp.x, __fdiv_rn(mbf, d));   // kpU.pt.x - mbf/d (src/Frame.cc:661)
            }
        }
    }
    uright[i] = ur;
    depth[i] = dz;
}

// Frame::UndistortKeyPoints (src/Frame.cc:404-434) -> cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK): normalise with
// K, 5 fixed-point iterations of the distortion model (k1 k2 p1 p2 k3 k4 k5 k6; the thin-prism terms are zero for ORB-SLAM2's
// 4 / 5 coefficient calibrations), re-project with K - all in double like OpenCV, result stored as float. The build uses
// --fmad=false, so no product is fused: identical to oracle/cvprim_mat.h (pinned to cv2 4.13) bit for bit.
struct UndistortParams { double fx, fy, cx, cy, ifx, ify, k[8]; };
__device__ __host__ inline void undistort_point(const UndistortParams& p, float xin, float yin, float* xo, float* yo) {
    const double x0 = ((double)xin - p.cx) * p.ifx, y0 = ((double)yin - p.cy) * p.ify;
    double x = x0, y = y0;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((p.k[7] * r2 + p.k[6]) * r2 + p.k[5]) * r2) / (1 + ((p.k[4] * r2 + p.k[1]) * r2 + p.k[0]) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dx = 2 * p.k[2] * x * y + p.k[3] * (r2 + 2 * x * x);
        const double dy = p.k[2] * (r2 + 2 * y * y) + 2 * p.k[3] * x * y;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    *xo = (float)(p.fx * x + p.cx);
    *yo = (float)(p.fy * y + p.cy);
}
__global__ void __launch_bounds__(256) undistort_keypoints_kernel(UndistortParams p, const orbx_keypoint* __restrict__ kps,
                                                                  const int* __restrict__ count, int cap, orbx_keypoint* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(*count, cap)) return;
    orbx_keypoint kp = kps[i];
    undistort_point(p, kp.x, kp.y, &kp.x, &kp.y);
    out[i] = kp;
}
static int make_undistort_params(float fx, float fy, float cx, float cy, const float* dist, int n_dist, UndistortParams* p) {
    ORB_REQUIRE(dist && (n_dist == 4 || n_dist == 5 || n_dist == 8), "distortion coefficients: 4 (k1 k2 p1 p2), 5 (+ k3) or 8 (+ k4 k5 k6)");
    ORB_REQUIRE(fx != 0 && fy != 0, "zero focal length");
    p->fx = fx; p->fy = fy; p->cx = cx; p->cy = cy;
    p->ifx = 1.0 / p->fx; p->ify = 1.0 / p->fy;
    for (int i = 0; i < 8; ++i) p->k[i] = i < n_dist ? (double)dist[i] : 0.0;
    return ORB_OK;
}

}  // namespace orb

extern "C" int orbm_undistort_keypoints_device(int device, const orbx_keypoint* d_kps, const int32_t* d_count, int cap, float fx, float fy, float cx,
                                               float cy, const float* dist_coef, int n_dist, orbx_keypoint* d_kps_un, void* stream) {
    using namespace orb;
    ORB_REQUIRE(d_kps && d_count && d_kps_un && cap > 0, "bad arguments");
    UndistortParams p;
    const int rc = make_undistort_params(fx, fy, cx, cy, dist_coef, n_dist, &p);
    if (rc != ORB_OK) return rc;
    ORB_CUDA_TRY(cudaSetDevice(device));
    if (dist_coef[0] == 0.0f) {   // mDistCoef.at<float>(0) == 0.0: mvKeysUn = mvKeys (src/Frame.cc:406-410)
        ORB_CUDA_TRY(cudaMemcpyAsync(d_kps_un, d_kps, (size_t)cap * sizeof(orbx_keypoint), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return ORB_OK;
    }
    undistort_keypoints_kernel<<<ceil_div(cap, 256), 256, 0, (cudaStream_t)stream>>>(p, d_kps, d_count, cap, d_kps_un);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// Frame::ComputeImageBounds (src/Frame.cc:436-464), host only: the four image corners through the same arithmetic.
extern "C" int orbm_image_bounds(int width, int height, float fx, float fy, float cx, float cy, const float* dist_coef, int n_dist, float* bounds4) {
    using namespace orb;
    ORB_REQUIRE(bounds4 && width > 0 && height > 0, "bad arguments");
    UndistortParams p;
    const int rc = make_undistort_params(fx, fy, cx, cy, dist_coef, n_dist, &p);
    if (rc != ORB_OK) return rc;
    if (dist_coef[0] == 0.0f) { bounds4[0] = 0.0f; bounds4[1] = (float)width; bounds4[2] = 0.0f; bounds4[3] = (float)height; return ORB_OK; }
    const float cxs[4] = {0.f, (float)width, 0.f, (float)width}, cys[4] = {0.f, 0.f, (float)height, (float)height};
    float ux[4], uy[4];
    for (int i = 0; i < 4; ++i) undistort_point(p, cxs[i], cys[i], &ux[i], &uy[i]);
    bounds4[0] = ux[0] < ux[2] ? ux[0] : ux[2];   // mnMinX = min(corner0.x, corner2.x)
    bounds4[1] = ux[1] > ux[3] ? ux[1] : ux[3];   // mnMaxX
    bounds4[2] = uy[0] < uy[1] ? uy[0] : uy[1];   // mnMinY
    bounds4[3] = uy[2] > uy[3] ? uy[2] : uy[3];   // mnMaxY
    return ORB_OK;
}

extern "C" int orbm_project_points_device(int device, const orbm_camera* cam, int mode, float viewing_cos_limit, float th,
                                          const float* d_world_pos, const float* d_normal, const float* d_max_distance,
                                          const float* d_min_distance, int n, uint8_t* d_alive, float* d_u, float* d_v,
                                          float* d_ur, int32_t* d_level, float* d_view_cos, float* d_radius,
                                          int32_t* d_min_level, int32_t* d_max_level, void* stream) {
    using namespace orb;
    ORB_REQUIRE(cam && n >= 0 && (mode == 0 || mode == 1), "bad arguments");
    ORB_REQUIRE(cam->n_levels >= 1 && cam->n_levels <= ORBM_MAX_LEVELS, "n_levels out of range");
    if (n == 0) return ORB_OK;
    ORB_REQUIRE(d_world_pos && d_normal && d_max_distance && d_min_distance && d_alive && d_u && d_v && d_ur && d_level && d_view_cos,
                "null pointer");
    ORB_CUDA_TRY(cudaSetDevice(device));
    ProjectParams p;
    p.cam = *cam;
    p.cos_limit = viewing_cos_limit;
    p.th = th;
    p.mode = mode;
    p.n = n;
    project_points_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p, d_world_pos, d_normal, d_max_distance, d_min_distance,
                                                                             d_alive, d_u, d_v, d_ur, d_level, d_view_cos, d_radius,
                                                                             d_min_level, d_max_level);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

extern "C" int orbm_stereo_from_rgbd_device(int device, const orbx_keypoint* d_kps, const orbx_keypoint* d_kps_un, const int32_t* d_count, int cap,
                                            const float* d_depth_image, int width, int height, size_t stride_bytes, float mbf, float* d_uRight,
                                            float* d_depth, void* stream) {
    using namespace orb;
    ORB_REQUIRE(d_kps && d_count && d_depth_image && d_uRight && d_depth, "null pointer");
    ORB_REQUIRE(cap > 0 && width > 0 && height > 0 && stride_bytes >= (size_t)width * 4 && stride_bytes % 4 == 0, "bad geometry");
    ORB_CUDA_TRY(cudaSetDevice(device));
    stereo_from_rgbd_kernel<<<ceil_div(cap, 256), 256, 0, (cudaStream_t)stream>>>(d_kps, d_kps_un, d_count, cap, d_depth_image, width, height,
                                                                                 stride_bytes / 4, mbf, d_uRight, d_depth);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}
