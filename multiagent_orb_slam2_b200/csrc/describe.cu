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
constexpr int kDescSlots = 32;  // keypoint slots per block

struct SlotInfo { int valid, level, x, y, response, dst; };

// grid (ceil(sel_words / 32), frames), block 256. Phases:
//   0. thread i < 32 resolves slot i (level, position in the frame's output, packed candidate)
//   A. every warp: intensity-centroid angle of 4 slots (lane = patch column, loop over rows)
//   B. warp 0: lane i evaluates sin/cos (double, rounded once) for slot i - 32 keypoints per
//      instruction stream instead of one
//   C. every warp: rotated BRIEF of 4 slots (lane = output byte), keypoint record
__global__ void __launch_bounds__(kDescWarps * 32, 5)
orient_describe_kernel(const Geometry* __restrict__ g, FrameSet fs, const uint8_t* __restrict__ pyr,
                       const uint8_t* __restrict__ blur, const uint32_t* __restrict__ selected,
                       const int* __restrict__ sel_counts, const float* __restrict__ pattern,
                       orbx_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int* __restrict__ counts) {
    __shared__ float patf[1024];  // transposed: value (test t, component c) of byte `lane` at [(4t+c)*32 + lane]
    __shared__ SlotInfo info[kDescSlots];
    __shared__ float s_angle[kDescSlots], s_cos[kDescSlots], s_sin[kDescSlots];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) patf[i] = pattern[i];

    const int frame = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nlevels = g->nlevels;
    const int* cnt = sel_counts + (size_t)frame * nlevels;
    if (blockIdx.x == 0 && threadIdx.x == 32) {
        int total = 0;
        for (int l = 0; l < nlevels; ++l) total += cnt[l];
        counts[frame] = min(total, g->out_cap);
    }
    if (threadIdx.x < kDescSlots) {
        const int slot = blockIdx.x * kDescSlots + threadIdx.x;
        SlotInfo si{0, 0, 0, 0, 0, 0};
        if (slot < g->sel_words) {
            int level = 0, dst0 = 0;
            for (int l = 0; l < nlevels; ++l) if (slot >= g->lv[l].sel_off) level = l;
            for (int l = 0; l < level; ++l) dst0 += cnt[l];
            const int i = slot - g->lv[level].sel_off;
            const int dst = dst0 + i;
            if (i < cnt[level] && dst < g->out_cap) {
                const uint32_t p = selected[(size_t)frame * g->sel_words + slot];
                si.valid = 1; si.level = level; si.dst = dst;
                si.x = (int)(p & 0xfff) + kMinBorder; si.y = (int)((p >> 12) & 0xfff) + kMinBorder; si.response = (int)(p >> 24);
            }
        }
        info[threadIdx.x] = si;
    }
    __syncthreads();

    // ---- A: IC_Angle -----------------------------------------------------------------------------
    for (int k = 0; k < kDescSlots / kDescWarps; ++k) {
        const int sidx = warp * (kDescSlots / kDescWarps) + k;
        const SlotInfo si = info[sidx];
        if (!si.valid) continue;
        int spitch;
        const uint8_t* img = level_ptr(*g, fs, pyr, frame, si.level, &spitch);
        const uint8_t* center = img + (size_t)si.y * spitch + si.x;
        int m10 = 0, m01 = 0;
        const int u = lane - kHalfPatch;
        if (lane < 2 * kHalfPatch + 1) {
            // the circular patch is symmetric under transposition: column u spans rows |v| <= umax[|u|]
            const int vext = g->umax[u < 0 ? -u : u];
            const uint8_t* col = center + u;
            int vals[2 * kHalfPatch + 1];
#pragma unroll
            for (int v = -kHalfPatch; v <= kHalfPatch; ++v) {
                const int av = v < 0 ? -v : v;
                vals[v + kHalfPatch] = av <= vext ? (int)col[v * spitch] : 0;  // all loads in flight together
            }
            int colsum = 0;
#pragma unroll
            for (int v = -kHalfPatch; v <= kHalfPatch; ++v) { colsum += vals[v + kHalfPatch]; m01 += v * vals[v + kHalfPatch]; }
            m10 = u * colsum;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
        if (lane == 0) s_angle[sidx] = fast_atan2_deg((float)m01, (float)m10);
    }
    __syncthreads();

    // ---- B: sin / cos, one keypoint per lane -------------------------------------------------------
    if (warp == 0 && info[lane].valid) {
        const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
        const float rad = __fmul_rn(s_angle[lane], factorPI);
        double sn, cs;
        sincos((double)rad, &sn, &cs);
        s_cos[lane] = (float)cs; s_sin[lane] = (float)sn;
    }
    __syncthreads();

    // ---- C: rotated BRIEF + keypoint record ----------------------------------------------------------
    for (int k = 0; k < kDescSlots / kDescWarps; ++k) {
        const int sidx = warp * (kDescSlots / kDescWarps) + k;
        const SlotInfo si = info[sidx];
        if (!si.valid) continue;
        const LevelGeom& L = g->lv[si.level];
        const float a = s_cos[sidx], b = s_sin[sidx];
        const uint8_t* bc = blur + (size_t)frame * g->blur_bytes + L.blur_off + (size_t)si.y * L.pitch + si.x;
        const int bp = L.pitch;
        const float* pp = patf + lane;
        int val = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const float x0 = pp[(4 * t) * 32], y0 = pp[(4 * t + 1) * 32], x1 = pp[(4 * t + 2) * 32], y1 = pp[(4 * t + 3) * 32];
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            val |= (int)(bc[r0 * bp + c0] < bc[r1 * bp + c1]) << t;
        }
        desc[((size_t)frame * g->out_cap + si.dst) * 32 + lane] = (uint8_t)val;
        if (lane == 0) {
            orbx_keypoint kp;
            // `pt *= scale` (1095-1101) is applied for level != 0 only; scale[0] == 1 makes it uniform
            kp.x = __fmul_rn((float)si.x, L.scale);
            kp.y = __fmul_rn((float)si.y, L.scale);
            kp.size = (float)L.patch_size;
            kp.angle = s_angle[sidx];
            kp.response = (float)si.response;
            kp.octave = si.level;
            kps[(size_t)frame * g->out_cap + si.dst] = kp;
        }
    }
}

int launch_describe(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int n, cudaStream_t st) {
    orient_describe_kernel<<<dim3(ceil_div(hg.sel_words, kDescSlots), n), kDescWarps * 32, 0, st>>>(
        db.geom, fs, db.pyr, db.blur, db.selected, db.sel_counts, db.pattern, db.kps, db.desc, db.counts);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
