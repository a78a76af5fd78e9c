// K9: stereo matching of a rectified pair, one warp per left keypoint.
//
// Replaces Frame::ComputeStereoMatches, /root/reference/src/Frame.cc:466-640:
//   * candidates of a left keypoint = right keypoints whose row band [floor(y-r), ceil(y+r)],
//     r = 2*scale[octave], contains int(vL) (the vRowIndices table, 477-495), |octave difference| <= 1,
//     uR in [uL - maxD, uL]; best Hamming distance below TH_HIGH, first (lowest index) minimum wins;
//   * if best < 75: 11 x 11 L1 patch distance (centre-subtracted) at 11 horizontal shifts on the left
//     keypoint's pyramid level, parabola sub-pixel fit, disparity gates (544-620);
//   * final rejection of matches whose patch distance is >= 1.5*1.4*median (625-639).
// The patch sums are integers (exactly what the reference's float sums hold), the parabola is
// evaluated with explicit round-to-nearest float ops: results are bit-exact against the oracle.
// The pyramids are read in place from the two extractor handles (device resident).
#include <climits>

#include "extract_kernels.cuh"

namespace orb {

constexpr int kStWarps = 8;

__device__ __forceinline__ int st_reflect(int p, int n) {
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * n - 2 - p : p;
    return max(0, min(p, n - 1));
}


__global__ void __launch_bounds__(kStWarps * 32)
stereo_match_kernel(StereoSide Ls, StereoSide Rs, float mbf, float mb, float* __restrict__ uRight, float* __restrict__ depth,
                    int* __restrict__ sad, int out_stride) {
    __shared__ uint8_t sL[kStWarps][128], sR[kStWarps][11 * 21 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * kStWarps + warp;
    // blockIdx.y = frame of the batch, relative to the first one (Ls.frame / Rs.frame)
    const int fr = blockIdx.y;
    Ls.frame += fr; Rs.frame += fr;
    Ls.kps += (size_t)fr * Ls.g->out_cap; Ls.desc += (size_t)fr * Ls.g->out_cap * 32; Ls.count += fr;
    Rs.kps += (size_t)fr * Rs.g->out_cap; Rs.desc += (size_t)fr * Rs.g->out_cap * 32; Rs.count += fr;
    uRight += (size_t)fr * out_stride; depth += (size_t)fr * out_stride; sad += (size_t)fr * out_stride;
    const int nL = min(*Ls.count, Ls.g->out_cap), nR = min(*Rs.count, Rs.g->out_cap);
    if (iL >= nL) return;
    const orbx_keypoint kL = Ls.kps[iL];
    const int levelL = kL.octave;
    const float uL = kL.x, vL = kL.y;
    const int row = (int)vL;
    const float maxD = __fdiv_rn(mbf, mb);
    const float minU = __fsub_rn(uL, maxD), maxU = uL;
    if (lane == 0) { uRight[iL] = -1.0f; depth[iL] = -1.0f; sad[iL] = -1; }
    if (maxU < 0) return;

    const uint4* dl = reinterpret_cast<const uint4*>(Ls.desc + (size_t)iL * 32);
    const uint4 a0 = __ldg(dl), a1 = __ldg(dl + 1);
    uint32_t best = (100u << 16) | 0xffffu;
    for (int iR = lane; iR < nR; iR += 32) {
        const orbx_keypoint kR = Rs.kps[iR];
        const float r = __fmul_rn(2.0f, Rs.g->lv[kR.octave].scale);
        const int maxr = (int)ceilf(__fadd_rn(kR.y, r)), minr = (int)floorf(__fsub_rn(kR.y, r));
        if (row < minr || row > maxr) continue;
        if (kR.octave < levelL - 1 || kR.octave > levelL + 1) continue;
        if (!(kR.x >= minU && kR.x <= maxU)) continue;
        const uint4* dr = reinterpret_cast<const uint4*>(Rs.desc + (size_t)iR * 32);
        const uint4 b0 = __ldg(dr), b1 = __ldg(dr + 1);
        const uint32_t d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
        best = min(best, (d << 16) | (uint32_t)iR);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((best >> 16) >= 75u) return;  // thOrbDist = (TH_HIGH + TH_LOW) / 2
    const int bestIdxR = (int)(best & 0xffffu);

    const float uR0 = Rs.kps[bestIdxR].x;
    const float sf = __fdiv_rn(1.0f, Ls.g->lv[levelL].scale);  // mvInvScaleFactors
    const float scaleduL = roundf(__fmul_rn(uL, sf)), scaledvL = roundf(__fmul_rn(vL, sf)), scaleduR0 = roundf(__fmul_rn(uR0, sf));
    const int wL = Ls.g->lv[levelL].w, hL = Ls.g->lv[levelL].h, wR = Rs.g->lv[levelL].w, hR = Rs.g->lv[levelL].h;
    if (scaleduR0 < 0 || __fadd_rn(scaleduR0, 11.0f) >= (float)wR) return;  // iniu / endu test (568-571)
    const int cuL = (int)scaleduL, cvL = (int)scaledvL, cuR = (int)scaleduR0;
    int pL, pR;
    const uint8_t* imL = level_ptr(*Ls.g, Ls.fs, Ls.pyr, Ls.frame, levelL, &pL);
    const uint8_t* imR = level_ptr(*Rs.g, Rs.fs, Rs.pyr, Rs.frame, levelL, &pR);
    // pixels outside the level = the reference's BORDER_REFLECT_101 frame
    for (int i = lane; i < 121; i += 32) {
        const int dy = i / 11 - 5, dx = i % 11 - 5;
        sL[warp][i] = imL[(size_t)st_reflect(cvL + dy, hL) * pL + st_reflect(cuL + dx, wL)];
    }
    for (int i = lane; i < 231; i += 32) {
        const int dy = i / 21 - 5, dx = i % 21 - 10;
        sR[warp][i] = imR[(size_t)st_reflect(cvL + dy, hR) * pR + st_reflect(cuR + dx, wR)];
    }
    __syncwarp();
    int dist = INT_MAX;
    if (lane < 11) {
        const int inc = lane - 5;
        const int cL = sL[warp][60], cR = sR[warp][5 * 21 + 10 + inc];
        int s = 0;
        for (int dy = 0; dy < 11; ++dy)
#pragma unroll
            for (int dx = 0; dx < 11; ++dx) s += abs((sL[warp][dy * 11 + dx] - cL) - (sR[warp][dy * 21 + dx + 5 + inc] - cR));
        dist = s;
    }
    // first minimum over ascending shifts (strict '<' updates, 589-593)
    uint32_t key = lane < 11 ? ((uint32_t)dist << 8) | (uint32_t)lane : 0xffffffffu;
#pragma unroll
    for (int o = 16; o; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    const int bl = (int)(key & 0xffu), bestDist = (int)(key >> 8);
    if (bl == 0 || bl == 10) return;
    const float d1 = (float)__shfl_sync(0xffffffffu, dist, bl - 1), d2 = (float)bestDist, d3 = (float)__shfl_sync(0xffffffffu, dist, bl + 1);
    if (lane != 0) return;
    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = __fmul_rn(Ls.g->lv[levelL].scale, __fadd_rn(__fadd_rn(scaleduR0, (float)(bl - 5)), deltaR));
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= 0 && disparity < maxD) {
        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }  // `uL-0.01` is double arithmetic (612-616)
        depth[iL] = __fdiv_rn(mbf, disparity);
        uRight[iL] = bestuR;
        sad[iL] = bestDist;
    }
}

// Median of the patch distances and the 1.5*1.4*median rejection (src/Frame.cc:625-639), one block per frame. The
// reference sorts (distance, index) pairs and takes element m/2: found here by a radix select over the distances (they are
// below 2^17: 121 pixels x 2 x 255) - 17 block-wide counting passes instead of n^2 rank comparisons - the distances staying
// in global memory (L2), so there is no shared-memory limit on the number of keypoints.
__global__ void __launch_bounds__(1024)
stereo_median_filter_kernel(const int* __restrict__ count, int cap, float* __restrict__ uRight, float* __restrict__ depth,
                            int* __restrict__ sad, int* __restrict__ kept, int out_stride) {
    __shared__ int s_cnt[2], s_m;
    const int fr = blockIdx.x;
    count += fr; kept += fr;
    uRight += (size_t)fr * out_stride; depth += (size_t)fr * out_stride; sad += (size_t)fr * out_stride;
    const int n = min(*count, cap);
    if (threadIdx.x == 0) s_m = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine += sad[i] >= 0;
    if (mine) atomicAdd(&s_m, mine);
    __syncthreads();
    const int m = s_m;
    if (m == 0) { if (threadIdx.x == 0) *kept = 0; return; }
    // the (m/2)-th smallest distance (0-based; equal distances are interchangeable for the value)
    int prefix = 0, k = m / 2;
    for (int bit = 17; bit >= 0; --bit) {
        __syncthreads();
        if (threadIdx.x == 0) s_cnt[0] = 0;
        __syncthreads();
        int zeros = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int d = sad[i];
            if (d >= 0 && (d >> (bit + 1)) == (prefix >> (bit + 1)) && !((d >> bit) & 1)) ++zeros;
        }
        if (zeros) atomicAdd(&s_cnt[0], zeros);
        __syncthreads();
        const int z = s_cnt[0];
        if (k >= z) { k -= z; prefix |= 1 << bit; }
    }
    const float thDist = __fmul_rn(1.5f * 1.4f, (float)prefix);
    int good = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int d = sad[i];
        if (d < 0) continue;
        if ((float)d < thDist) ++good;
        else { uRight[i] = -1.0f; depth[i] = -1.0f; }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_m = 0;
    __syncthreads();
    if (good) atomicAdd(&s_m, good);
    __syncthreads();
    if (threadIdx.x == 0) *kept = s_m;
}

// n_frames consecutive frames of both handles, starting at L.frame / R.frame; outputs [n_frames][out_stride]
int launch_stereo(const StereoSide& L, const StereoSide& R, int capL, int n_frames, int out_stride, float mbf, float mb, float* d_uRight,
                  float* d_depth, int* d_sad, int* d_kept, cudaStream_t st) {
    ORB_REQUIRE(n_frames >= 1 && n_frames <= 65535, "1..65535 frames per stereo call");
    stereo_match_kernel<<<dim3(ceil_div(capL, kStWarps), n_frames), kStWarps * 32, 0, st>>>(L, R, mbf, mb, d_uRight, d_depth, d_sad, out_stride);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    stereo_median_filter_kernel<<<n_frames, 1024, 0, st>>>(L.count, capL, d_uRight, d_depth, d_sad, d_kept, out_stride);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
