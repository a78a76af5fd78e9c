// K5 + K6: orientation (intensity centroid) and rotated-BRIEF descriptor, one warp per keypoint.
//
// Replaces IC_Angle (/root/reference/src/ORBextractor.cc:77-104), computeOrbDescriptor (108-147) and
// the output assembly of operator() (1075-1104): keypoints of level l land at
// sum_{k<l} count_k + i in the frame's result arrays, coordinates scaled by mvScaleFactor[l].
//
// Float discipline (bit-exact against the oracle):
//   * moments are int32 (exact), cv::fastAtan2's degree-7 polynomial is evaluated with explicit
//     round-to-nearest mul/add/div (no FMA contraction);
//   * cos/sin of the float angle are evaluated in double and rounded once (== glibc cosf/sinf up
//     to their rare non-correctly-rounded cases), the rotation x*b + y*a uses __fmul_rn/__fadd_rn,
//     cvRound is round-half-even (__float2int_rn).
#include <cfloat>

#include "extract_kernels.cuh"

namespace orb {

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)DBL_EPSILON;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int kDescWarps = 8;

// grid (ceil(sel_words / 8), frames), block 256 = 8 keypoint slots
__global__ void __launch_bounds__(kDescWarps * 32)
orient_describe_kernel(const Geometry* __restrict__ g, FrameSet fs, const uint8_t* __restrict__ pyr,
                       const uint8_t* __restrict__ blur, const uint32_t* __restrict__ selected,
                       const int* __restrict__ sel_counts, const int8_t* __restrict__ pattern,
                       orbx_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int* __restrict__ counts) {
    __shared__ int8_t pat[1024];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) reinterpret_cast<int*>(pat)[i] = reinterpret_cast<const int*>(pattern)[i];
    __syncthreads();

    const int frame = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * kDescWarps + warp;
    const int nlevels = g->nlevels;
    const int* cnt = sel_counts + (size_t)frame * nlevels;
    if (slot == 0 && lane == 0) {
        int total = 0;
        for (int l = 0; l < nlevels; ++l) total += cnt[l];
        counts[frame] = total;
    }
    if (slot >= g->sel_words) return;
    // which level does this slot belong to, and where does the level start in the output?
    int level = 0, dst0 = 0;
    for (int l = 0; l < nlevels; ++l) {
        if (slot >= g->lv[l].sel_off) level = l;
    }
    for (int l = 0; l < level; ++l) dst0 += cnt[l];
    const LevelGeom& L = g->lv[level];
    const int i = slot - L.sel_off;
    if (i >= cnt[level]) return;
    const int dst = dst0 + i;
    if (dst >= g->out_cap) return;

    const uint32_t p = selected[(size_t)frame * g->sel_words + slot];
    const int x = (int)(p & 0xfff) + kMinBorder, y = (int)((p >> 12) & 0xfff) + kMinBorder;
    const int response = (int)(p >> 24);

    // ---- IC_Angle: lane = column u in [-15, 15], loop over rows ---------------------------------
    int spitch;
    const uint8_t* img = level_ptr(*g, fs, pyr, frame, level, &spitch);
    const uint8_t* center = img + (size_t)y * spitch + x;
    int m10 = 0, m01 = 0;
    const int u = lane - kHalfPatch;
    if (lane < 2 * kHalfPatch + 1) {
        const int au = u < 0 ? -u : u;
#pragma unroll 4
        for (int v = -kHalfPatch; v <= kHalfPatch; ++v) {
            const int av = v < 0 ? -v : v;
            if (au <= g->umax[av]) {
                const int val = center[v * spitch + u];
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- rotated BRIEF: lane = output byte, 8 tests x 2 samples ---------------------------------
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    const float a = (float)cos((double)rad), b = (float)sin((double)rad);
    const uint8_t* bc = blur + (size_t)frame * g->blur_bytes + L.blur_off + (size_t)y * L.pitch + x;
    const int bp = L.pitch;
    const int8_t* pp = pat + lane * 32;
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float x0 = (float)pp[4 * t], y0 = (float)pp[4 * t + 1], x1 = (float)pp[4 * t + 2], y1 = (float)pp[4 * t + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        val |= (int)(bc[r0 * bp + c0] < bc[r1 * bp + c1]) << t;
    }
    desc[((size_t)frame * g->out_cap + dst) * 32 + lane] = (uint8_t)val;

    if (lane == 0) {
        orbx_keypoint k;
        // `pt *= scale` (1095-1101) is applied for level != 0 only; scale[0] == 1 makes it uniform
        k.x = __fmul_rn((float)x, L.scale);
        k.y = __fmul_rn((float)y, L.scale);
        k.size = (float)L.patch_size;
        k.angle = angle;
        k.response = (float)response;
        k.octave = level;
        kps[(size_t)frame * g->out_cap + dst] = k;
    }
}

int launch_describe(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int n, cudaStream_t st) {
    orient_describe_kernel<<<dim3(ceil_div(hg.sel_words, kDescWarps), n), kDescWarps * 32, 0, st>>>(
        db.geom, fs, db.pyr, db.blur, db.selected, db.sel_counts, db.pattern, db.kps, db.desc, db.counts);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
