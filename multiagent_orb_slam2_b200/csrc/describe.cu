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
constexpr int kDescSlotsMax = 32;  // keypoint slots per block: 8 warps x SPW (4 for batches, 1 for one or two frames: four
                                   // times the blocks, a warp's keypoints no longer queue behind each other)
// blurred window of a keypoint in shared memory: rows y-18..y+18, 40 bytes from the word at or left of x-18
// (64-bit loads of a 48-byte window were tried: 43 KB of shared memory per block pushes the SM to its largest
// carve-out and the kernel, which also lives on L1 hits, from 0.49 to 0.64 ms; cudaFuncAttributePreferredSharedMemoryCarveout
// swept 40..100 %: the default choice is the best)
constexpr int kPatchReach = 18, kPatchRows = 2 * kPatchReach + 1, kPatchRowWords = 10, kPatchWords = 372;

struct SlotInfo { int valid, level, x, y, response, dst; };

// grid (ceil(sel_words / 32), frames), block 256. Phases:
//   0. thread i < 32 resolves slot i (level, position in the frame's output, packed candidate)
//   A. every warp: intensity-centroid moments of 4 slots (aligned word loads, IDP.4A against int8 coordinates)
//   B. every warp: lanes 0-3 evaluate the angle and sin/cos (double, rounded once) of the warp's own 4 slots;
//      after the table fill no block barrier separates the phases
//   C. every warp: rotated BRIEF of 4 slots, two at a time, sampled from a shared-memory copy of the
//      blurred 37 x 40 window (lane = output byte), keypoint record
template <int SPW>
__global__ void __launch_bounds__(kDescWarps * 32, 5)
orient_describe_kernel(const Geometry* __restrict__ g, FrameSet fs, const uint8_t* __restrict__ pyr,
                       const uint8_t* __restrict__ blur, const uint32_t* __restrict__ selected,
                       const int* __restrict__ sel_counts, const uint32_t* __restrict__ pattern, const uint32_t* __restrict__ ic_table,
                       orbx_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int* __restrict__ counts) {
    // Shared memory is kept below 31.8 KB per block on purpose: five blocks then fit the 164 KB carve-out and leave 92 KB of L1
    // to the window gathers (37.7 KB per block = 196 KB carve-out measured 0.469 ms, this layout 0.36 ms at B=512).
    __shared__ uint32_t pat8[kPatternWords];  // x0 | y0 << 8 | x1 << 16 | y1 << 24 (int8) of test t of byte `lane` at [t * 32 + lane]
    constexpr int kDescSlots = kDescWarps * SPW;
    __shared__ SlotInfo info[kDescSlots + 1];   // + 1: the pair partner of the last slot when SPW is odd (always invalid)
    __shared__ float s_angle[kDescSlots + 1], s_cos[kDescSlots + 1], s_sin[kDescSlots + 1];
    __shared__ __align__(16) uint32_t patch[kDescWarps * 2 * kPatchWords];  // per warp: blurred windows of two keypoints
    __shared__ int s_m10[kDescSlots], s_m01[kDescSlots];
    if (threadIdx.x == 0) info[kDescSlots] = SlotInfo{0, 0, 0, 0, 0, 0};
    __shared__ uint32_t ictab[kIcTableWords];  // [phase][item]: u + 16 per byte, 0 outside the patch
    for (int i = threadIdx.x; i < kPatternWords; i += blockDim.x) pat8[i] = pattern[i];
    for (int i = threadIdx.x; i < kIcTableWords; i += blockDim.x) ictab[i] = ic_table[i];
    pdl_wait();   // programmatic dependent of the quadtree launch: the tables above are filled while the trees are still built

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

    constexpr int kIcRows = 2 * kHalfPatch + 1;
    const int ic_lr = lane / kIcWordsPerRow, ic_lj = lane - ic_lr * kIcWordsPerRow;
    // ---- A: IC_Angle moments ---------------------------------------------------------------------
    // The 31 x 31 bounding box of the circular patch is read as aligned 32-bit words (9 per row, the left
    // edge rounded down to a word), lane = (row, word) item; the table holds, per alignment phase and
    // item, the u and v coordinates of the 4 bytes as int8 (0 outside the circle), so that two IDP.4A
    // per word accumulate m10 = sum u I and m01 = sum v I (integer sums: same value as the reference's
    // row-pair loop, whatever the order).
    for (int k = 0; k < kDescSlots / kDescWarps; ++k) {
        const int sidx = warp * (kDescSlots / kDescWarps) + k;
        const SlotInfo si = info[sidx];
        if (!si.valid) continue;
        int spitch;
        const uint8_t* img = level_ptr(*g, fs, pyr, frame, si.level, &spitch);
        const int xs = si.x - kHalfPatch;
        const uint8_t* base = img + (size_t)(si.y - kHalfPatch) * spitch + (xs & ~3);
        const uint32_t* tu = ictab + (xs & 3) * kIcItems + lane;
        // 3 rows (27 words) per warp step: item 27 s + lane, its word at a fixed per-lane offset + s * 3 rows
        const uint8_t* wp = base + ic_lr * spitch + 4 * ic_lj;
        const int step = 3 * spitch;
        // all loads first (unconditional: the idle lanes' addresses are inside the image too), so that they
        // are in flight together
        constexpr int kSteps = (kIcRows * kIcWordsPerRow + 26) / 27;
        uint32_t w[kSteps];
#pragma unroll
        for (int s = 0; s < kSteps; ++s) w[s] = *reinterpret_cast<const uint32_t*>(wp + s * step);
        // per word: a = sum (u + 16) I and n = sum I over its bytes inside the patch ("inside" <=> table byte != 0);
        // m10 = sum a - 16 sum n, m01 = sum over rows of v * n (v is the word's row: 3 s + ic_lr - 15)
        uint32_t acc_a = 0, acc_n = 0;
        int acc_v = 0;
#pragma unroll
        for (int s = 0; s < kSteps; ++s) {
            const bool on = lane < 27 && 27 * s + lane < kIcRows * kIcWordsPerRow;
            const uint32_t cu = on ? tu[27 * s] : 0u;
            const uint32_t inside = ((cu + 0x7f7f7f7fu) >> 7) & 0x01010101u;   // bytes are <= 31: no carry between them
            const uint32_t n = __dp4a(w[s], inside, 0u);
            acc_a = __dp4a(w[s], cu, acc_a);
            acc_n += n;
            acc_v += (3 * s + ic_lr - kHalfPatch) * (int)n;
        }
        int m10 = (int)acc_a - 16 * (int)acc_n, m01 = acc_v;
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        if (lane == 0) { s_m10[sidx] = m10; s_m01[sidx] = m01; }
    }
    __syncwarp();

    // ---- B: angle, sin / cos: every warp for its own slots, one keypoint per lane (no block barrier between
    // the phases: a warp runs A, B, C for its 4 keypoints on its own) ------------------------------------
    if (lane < kDescSlots / kDescWarps) {
        const int sl = warp * (kDescSlots / kDescWarps) + lane;
        if (info[sl].valid) {
            const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
            const float ang = fast_atan2_deg((float)s_m01[sl], (float)s_m10[sl]);
            const float rad = __fmul_rn(ang, factorPI);
            double sn, cs;
            sincos((double)rad, &sn, &cs);
            s_angle[sl] = ang; s_cos[sl] = (float)cs; s_sin[sl] = (float)sn;
        }
    }
    __syncwarp();

    // ---- C: rotated BRIEF + keypoint record ----------------------------------------------------------
    // The 512 sample points of a keypoint lie within 18 px of it (pattern radius 18.4). Gathering them
    // straight from global memory costs one L1 wavefront per image row touched by each of the 16 load
    // instructions (~11 each: the kernel was bound by the L1 data pipe). Instead the warp copies the
    // 37 x 40 byte window (left edge rounded down to a word) into shared memory with coalesced word
    // loads and gathers there. Two keypoints are described per pass so that the pattern values are read
    // once for both.
    uint32_t* mypatch = patch + warp * 2 * kPatchWords;
    const int pt_lr = lane / kPatchRowWords, pt_lj = lane - pt_lr * kPatchRowWords;
    for (int kk = 0; kk < kDescSlots / kDescWarps; kk += 2) {
        const int sidx0 = warp * (kDescSlots / kDescWarps) + kk;
        const SlotInfo si0 = info[sidx0], si1 = info[kk + 1 < SPW ? sidx0 + 1 : kDescSlots];
        if (!si0.valid && !si1.valid) continue;
        // (an invalid slot of the pair samples stale data around the window centre and stores nothing)
        int off[2] = {kPatchReach * kPatchRowWords * 4 + kPatchReach, kPatchReach * kPatchRowWords * 4 + kPatchReach};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const SlotInfo& si = q ? si1 : si0;
            if (!si.valid) continue;
            const LevelGeom& L = g->lv[si.level];
            const int xs = si.x - kPatchReach;
            const uint8_t* src = blur + (size_t)frame * g->blur_bytes + L.blur_off + (size_t)(si.y - kPatchReach) * L.pitch + (xs & ~3);
            // 3 rows (30 words) per warp step: word 30 s + lane of the window, at a fixed per-lane offset + s * 3 rows
            const uint8_t* wp = src + pt_lr * L.pitch + 4 * pt_lj;
            const int step = 3 * L.pitch;
            uint32_t* d = mypatch + q * kPatchWords + lane;
            constexpr int kSteps = (kPatchRows * kPatchRowWords + 29) / 30;
            uint32_t w[kSteps];  // unconditional loads, all in flight
#pragma unroll
            for (int s = 0; s < kSteps; ++s)  // the last step's rows are clamped to the window: rows below it may lie outside the slab
                w[s] = *reinterpret_cast<const uint32_t*>(s + 1 < kSteps ? wp + s * step : src + min(3 * s + pt_lr, kPatchRows - 1) * L.pitch + 4 * pt_lj);
#pragma unroll
            for (int s = 0; s < kSteps; ++s)
                if (lane < 30 && 30 * s + lane < kPatchRows * kPatchRowWords) d[30 * s] = w[s];
            off[q] = kPatchReach * (kPatchRowWords * 4) + kPatchReach + (xs & 3);
        }
        __syncwarp();
        const uint8_t* p0 = reinterpret_cast<const uint8_t*>(mypatch) + off[0];
        const uint8_t* p1 = reinterpret_cast<const uint8_t*>(mypatch + kPatchWords) + off[1];
        const float a0 = si0.valid ? s_cos[sidx0] : 0.f, b0 = si0.valid ? s_sin[sidx0] : 0.f;  // s_cos / s_sin are unset for invalid slots
        const float a1 = si1.valid ? s_cos[sidx0 + 1] : 0.f, b1 = si1.valid ? s_sin[sidx0 + 1] : 0.f;  // (valid only when kk + 1 < SPW)
        const uint32_t* pp = pat8 + lane;
        uint32_t val0 = 0, val1 = 0;  // bits enter at the bottom, test 7 first: sign of I(p0) - I(p1) by one funnel shift
        constexpr int kRow = kPatchRowWords * 4;
#pragma unroll
        for (int t = 7; t >= 0; --t) {
            const uint32_t pw = pp[t * 32];   // small integers: the conversions are exact
            const float x0 = (float)(int8_t)(pw & 0xff), y0 = (float)(int8_t)((pw >> 8) & 0xff), x1 = (float)(int8_t)((pw >> 16) & 0xff),
                        y1 = (float)(int8_t)(pw >> 24);
            {
                const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b0), __fmul_rn(y0, a0)));
                const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a0), __fmul_rn(y0, b0)));
                const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b0), __fmul_rn(y1, a0)));
                const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a0), __fmul_rn(y1, b0)));
                val0 = __funnelshift_l((uint32_t)((int)p0[r0 * kRow + c0] - (int)p0[r1 * kRow + c1]), val0, 1);
            }
            {
                const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b1), __fmul_rn(y0, a1)));
                const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a1), __fmul_rn(y0, b1)));
                const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b1), __fmul_rn(y1, a1)));
                const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a1), __fmul_rn(y1, b1)));
                val1 = __funnelshift_l((uint32_t)((int)p1[r0 * kRow + c0] - (int)p1[r1 * kRow + c1]), val1, 1);
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const SlotInfo& si = q ? si1 : si0;
            if (!si.valid) continue;
            const LevelGeom& L = g->lv[si.level];
            desc[((size_t)frame * g->out_cap + si.dst) * 32 + lane] = (uint8_t)(q ? val1 : val0);
            if (lane == 0) {
                orbx_keypoint kp;
                // `pt *= scale` (1095-1101) is applied for level != 0 only; scale[0] == 1 makes it uniform
                kp.x = __fmul_rn((float)si.x, L.scale);
                kp.y = __fmul_rn((float)si.y, L.scale);
                kp.size = (float)L.patch_size;
                kp.angle = s_angle[sidx0 + q];
                kp.response = (float)si.response;
                kp.octave = si.level;
                kps[(size_t)frame * g->out_cap + si.dst] = kp;
            }
        }
        __syncwarp();  // the next pair overwrites the patches
    }
}

int launch_describe(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int n, cudaStream_t st, bool pdl) {
    const bool spread = n <= 2;   // one keypoint per warp: the GPU is far from full, latency is what counts
    const dim3 grid(ceil_div(hg.sel_words, kDescWarps * (spread ? 1 : 4)), n);
    auto kernel = spread ? orient_describe_kernel<1> : orient_describe_kernel<4>;
    if (pdl)
        ORB_CUDA_TRY(launch_pdl(kernel, grid, dim3(kDescWarps * 32), 0, st, db.geom, fs, db.pyr, db.blur, db.selected, db.sel_counts, db.pattern,
                                db.ic_table, db.kps, db.desc, db.counts));
    else
        kernel<<<grid, kDescWarps * 32, 0, st>>>(db.geom, fs, db.pyr, db.blur, db.selected, db.sel_counts, db.pattern, db.ic_table, db.kps, db.desc,
                                                 db.counts);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
