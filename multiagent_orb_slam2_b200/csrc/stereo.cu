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
                    int* __restrict__ sad) {
    __shared__ uint8_t sL[kStWarps][128], sR[kStWarps][11 * 21 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * kStWarps + warp;
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

// single block: median of (sad, index) pairs by rank counting, then the 1.5*1.4*median rejection
__global__ void __launch_bounds__(1024)
stereo_median_filter_kernel(const int* __restrict__ count, int cap, float* __restrict__ uRight, float* __restrict__ depth,
                            int* __restrict__ sad, int* __restrict__ kept) {
    extern __shared__ int s_sad[];  // [cap]
    __shared__ int s_m, s_median;
    const int n = min(*count, cap);
    if (threadIdx.x == 0) { s_m = 0; s_median = -1; }
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_sad[i] = sad[i];
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine += s_sad[i] >= 0;
    if (mine) atomicAdd(&s_m, mine);
    __syncthreads();
    const int m = s_m;
    if (m == 0) { if (threadIdx.x == 0) *kept = 0; return; }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int d = s_sad[i];
        if (d < 0) continue;
        int rank = 0;
        for (int j = 0; j < n; ++j) { const int e = s_sad[j]; rank += e >= 0 && (e < d || (e == d && j < i)); }
        if (rank == m / 2) s_median = d;
    }
    __syncthreads();
    const float thDist = __fmul_rn(1.5f * 1.4f, (float)s_median);
    int k = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int d = s_sad[i];
        if (d < 0) continue;
        if ((float)d < thDist) ++k;
        else { uRight[i] = -1.0f; depth[i] = -1.0f; }
    }
    if (threadIdx.x == 0) s_m = 0;
    __syncthreads();
    if (k) atomicAdd(&s_m, k);
    __syncthreads();
    if (threadIdx.x == 0) *kept = s_m;
}

int launch_stereo(const StereoSide& L, const StereoSide& R, int capL, float mbf, float mb, float* d_uRight, float* d_depth, int* d_sad,
                  int* d_kept, cudaStream_t st) {
    stereo_match_kernel<<<ceil_div(capL, kStWarps), kStWarps * 32, 0, st>>>(L, R, mbf, mb, d_uRight, d_depth, d_sad);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    stereo_median_filter_kernel<<<1, 1024, (size_t)capL * sizeof(int), st>>>(L.count, capL, d_uRight, d_depth, d_sad, d_kept);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
