// K4: quadtree keypoint distribution, one thread block per (level, frame).
//
// Replaces ORBextractor::DistributeOctTree + ExtractorNode::DivideNode,
// /root/reference/src/ORBextractor.cc:481-763, which walk a std::list of nodes one at a time. The
// same result (same keypoints, same output order, canonical tie-break) is produced here by:
//   1. path keys: root index, then 2 bits per depth (bit0 right half, bit1 bottom half) from the
//      recursive ceil-halving of DivideNode (483-509) - independent per candidate;
//   2. a block radix sort by key, after which every tree node is a contiguous range + a depth;
//   3. a node-parallel simulation of the list: one step processes the expandable nodes in a given
//      order (list order for a full pass 606-665; (size desc, list position asc) in the final phase
//      676-737, the canonical form of the reference's sort of pair<int,Node*> at 684), pushes the
//      children of each to the front and stops once the list holds >= N nodes (730-731);
//   4. per surviving node the candidate with the highest response, earliest candidate on ties
//      (744-760).
// The CPU statement of exactly this formulation is oracle/quadtree_arrayform.cc (checked against the
// direct std::list restatement); integer only, so results are bit-exact.
#include <atomic>

#include "extract_kernels.cuh"

namespace orb {

// -DORB_QT_PROFILE (experiments only, ORB_EXTRA_NVCC_FLAGS): block (0, 0) prints clock64 marks of its phases
#ifdef ORB_QT_PROFILE
#include <cstdio>
__device__ long long g_qt_marks[64];
__device__ int g_qt_nmarks;
#define QT_MARK(tag) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && g_qt_nmarks < 62) { g_qt_marks[g_qt_nmarks] = clock64(); g_qt_tags[g_qt_nmarks++] = (tag); } } while (0)
__device__ int g_qt_tags[64];
#else
#define QT_MARK(tag) do { } while (0)
#endif

constexpr int kQtThreads = 256;  // several blocks per SM hide each other's barrier stalls
constexpr int kQtWarps = kQtThreads / 32;
constexpr int kQtSmemKeys = 4096;  // candidates per (level, frame) served from shared memory; more fall back to global

struct QtShared {
    int warp_tmp[32];
    int carry;
    int bcast[4];
    int wcount[kQtWarps][256];
};

// exclusive scan of in[0..len) into out[0..len) (may alias); returns the total to every thread.
// Short arrays (the node lists: a few hundred entries) are scanned by warp 0 alone - each lane owns a
// contiguous chunk - so that the whole scan costs two block barriers instead of six.
__device__ int block_exclusive_scan(const int* in, int* out, int len, QtShared& sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // inputs written by other warps are visible
    if (len <= kQtThreads) {  // the node lists: one element per thread, two short parallel stages
        const int v = (int)threadIdx.x < len ? in[threadIdx.x] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sh.warp_tmp[warp] = inc;
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kQtWarps; ++w) { const int t = sh.warp_tmp[w]; total += t; off += w < warp ? t : 0; }
        if ((int)threadIdx.x < len) out[threadIdx.x] = off + inc - v;
        __syncthreads();  // out[] visible; warp_tmp free for the next scan
        return total;
    }
    if (len <= 2 * kQtThreads) {  // two consecutive elements per thread (the cell lists of the largest level)
        const int i0 = 2 * (int)threadIdx.x;
        const int v0 = i0 < len ? in[i0] : 0, v1 = i0 + 1 < len ? in[i0 + 1] : 0, v = v0 + v1;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sh.warp_tmp[warp] = inc;
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kQtWarps; ++w) { const int t = sh.warp_tmp[w]; total += t; off += w < warp ? t : 0; }
        if (i0 < len) out[i0] = off + inc - v;
        if (i0 + 1 < len) out[i0 + 1] = off + inc - v + v0;
        __syncthreads();  // out[] visible; warp_tmp free for the next scan
        return total;
    }
    if (len <= 32 * 32) {
        if (warp == 0) {
            const int per = (len + 31) >> 5, lo = lane * per, hi = min(lo + per, len);
            int sum = 0;
            for (int i = lo; i < hi; ++i) sum += in[i];
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            int run = inc - sum;
            for (int i = lo; i < hi; ++i) { const int v = in[i]; out[i] = run; run += v; }
            if (lane == 31) sh.carry = inc;
        }
        __syncthreads();
        const int total = len > 0 ? sh.carry : 0;
        __syncthreads();  // nobody may still be reading carry when the next scan overwrites it
        return total;
    }
    if (threadIdx.x == 0) sh.carry = 0;
    __syncthreads();
    for (int base = 0; base < len; base += kQtThreads) {
        const int i = base + threadIdx.x;
        const int v = i < len ? in[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sh.warp_tmp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int w = lane < kQtWarps ? sh.warp_tmp[lane] : 0;
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
            sh.warp_tmp[lane] = winc - w;  // exclusive warp offsets
            if (lane == 31) sh.bcast[0] = winc;  // chunk total
        }
        __syncthreads();
        const int carry = sh.carry;
        if (i < len) out[i] = carry + sh.warp_tmp[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) sh.carry = carry + sh.bcast[0];
        __syncthreads();
    }
    const int total = sh.carry;
    __syncthreads();  // nobody may still be reading carry when the next scan resets it
    return total;
}

__device__ __forceinline__ uint32_t qt_path_key(int x, int y, int H, float rootW, int depth) {
    const int r = (int)__fdiv_rn((float)x, rootW);
    int x0 = (int)__fmul_rn(rootW, (float)r), x1 = (int)__fmul_rn(rootW, (float)(r + 1));
    int y0 = 0, y1 = H;
    uint32_t key = (uint32_t)r;
    for (int d = 0; d < depth; ++d) {
        const int xm = x0 + ((x1 - x0 + 1) >> 1), ym = y0 + ((y1 - y0 + 1) >> 1);
        const uint32_t xb = x >= xm, yb = y >= ym;
        x0 = xb ? xm : x0; x1 = xb ? x1 : xm;
        y0 = yb ? ym : y0; y1 = yb ? y1 : ym;
        key = (key << 2) | (yb << 1) | xb;
    }
    return key;
}

// LSD radix sort of (key, val) pairs, 8 bits per pass; result ends in (kA, vA) or (kB, vB): returns
// 0 / 1 accordingly. Buffers are global (L2 resident) so any candidate count fits.
// Every warp owns one contiguous chunk of the input and walks it twice per pass (count, scatter) with
// warp-level synchronisation only; the block meets 4 times per pass (the earlier tile-by-tile version
// met 3 times per 256 elements). Stable: chunks are ordered by warp, a chunk is walked in order.
__device__ int block_radix_sort(uint32_t* kA, uint32_t* vA, uint32_t* kB, uint32_t* vB, int n, int key_bits, QtShared& sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = ((n + kQtThreads - 1) / kQtThreads) * 32;  // elements per warp, a multiple of 32
    const int lo = min(warp * per, n), hi = min(lo + per, n);
    const uint32_t lt = (1u << lane) - 1u;
    int flip = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        const uint32_t* ki = flip ? kB : kA; const uint32_t* vi = flip ? vB : vA;
        uint32_t* ko = flip ? kA : kB;       uint32_t* vo = flip ? vA : vB;
        for (int i = threadIdx.x; i < kQtWarps * 256; i += kQtThreads) (&sh.wcount[0][0])[i] = 0;
        __syncthreads();
        // digit counts of this warp's chunk (no order needed: one shared-memory atomic per element)
        for (int i = lo + lane; i < hi; i += 32) atomicAdd(&sh.wcount[warp][(ki[i] >> shift) & 255u], 1);
        __syncthreads();
        // thread d: total of digit d, exclusive scan over the digits, then the start of every warp's run
        {
            int c[kQtWarps], v = 0;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) { c[w] = sh.wcount[w][threadIdx.x]; v += c[w]; }
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) sh.warp_tmp[warp] = inc;
            __syncthreads();
            int run = inc - v;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) run += w < warp ? sh.warp_tmp[w] : 0;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) { sh.wcount[w][threadIdx.x] = run; run += c[w]; }
        }
        __syncthreads();
        // scatter: wcount[warp][digit] is the running cursor of this warp's run of that digit
        for (int base = lo; base < hi; base += 32) {
            const int i = base + lane;
            const bool valid = i < hi;
            uint32_t k = 0, v = 0;
            if (valid) { k = ki[i]; v = vi[i]; }
            const uint32_t digit = valid ? (k >> shift) & 255u : 0x100u + lane;
            const uint32_t peers = __match_any_sync(0xffffffffu, digit);
            const int rank = __popc(peers & lt);
            int pos = 0;
            if (valid) pos = sh.wcount[warp][digit] + rank;
            __syncwarp();
            if (valid && rank == 0) sh.wcount[warp][digit] += __popc(peers);
            if (valid) { ko[pos] = k; vo[pos] = v; }
            __syncwarp();
        }
        flip ^= 1;
        __syncthreads();
    }
    return flip;
}

// The same stable LSD sort for sets that fit BOTH buffer pairs into shared memory (n <= kQtSmemKeys / 2: every level of a
// 1000-feature extraction). A warp's chunk (at most 8 x 32 elements) is read into registers once per pass and serves the
// counting and the scatter sweep; the match masks are kept too. No global memory, no load latency inside the sweeps.
constexpr int kQtRegSteps = 8;
__device__ int block_radix_sort_small(uint32_t* kA, uint32_t* vA, uint32_t* kB, uint32_t* vB, int n, int key_bits, QtShared& sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = ((n + kQtThreads - 1) / kQtThreads) * 32;  // elements per warp, a multiple of 32, <= 32 * kQtRegSteps
    const int lo = min(warp * per, n), hi = min(lo + per, n);
    const uint32_t lt = (1u << lane) - 1u;
    int flip = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        const uint32_t* ki = flip ? kB : kA; const uint32_t* vi = flip ? vB : vA;
        uint32_t* ko = flip ? kA : kB;       uint32_t* vo = flip ? vA : vB;
        for (int i = threadIdx.x; i < kQtWarps * 256; i += kQtThreads) (&sh.wcount[0][0])[i] = 0;
        uint32_t k[kQtRegSteps], peers[kQtRegSteps];   // (the values are re-read in the scatter sweep: registers are what caps the blocks per SM)
#pragma unroll
        for (int j = 0; j < kQtRegSteps; ++j) {
            const int i = lo + 32 * j + lane;
            k[j] = i < hi ? ki[i] : 0u;
            peers[j] = 0u;
        }
        __syncthreads();
        // lanes of a step with the same digit; the matches of all steps are independent, the counter updates behind them are not
#pragma unroll
        for (int j = 0; j < kQtRegSteps; ++j) {
            if (lo + 32 * j < hi) {  // warp-uniform
                const bool valid = lo + 32 * j + lane < hi;
                peers[j] = __match_any_sync(0xffffffffu, valid ? (k[j] >> shift) & 255u : 0x100u + lane);  // invalid lanes never group
            }
        }
        // (counting needs no order: one shared-memory atomic per element; the matches are only used by the scatter sweep and
        // are in flight meanwhile)
#pragma unroll
        for (int j = 0; j < kQtRegSteps; ++j)
            if (lo + 32 * j + lane < hi) atomicAdd(&sh.wcount[warp][(k[j] >> shift) & 255u], 1);
        __syncthreads();
        {
            int c[kQtWarps], t = 0;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) { c[w] = sh.wcount[w][threadIdx.x]; t += c[w]; }
            int inc = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            if (lane == 31) sh.warp_tmp[warp] = inc;
            __syncthreads();
            int run = inc - t;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) run += w < warp ? sh.warp_tmp[w] : 0;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) { sh.wcount[w][threadIdx.x] = run; run += c[w]; }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kQtRegSteps; ++j) {
            if (lo + 32 * j < hi) {
                const bool valid = lo + 32 * j + lane < hi;
                const uint32_t digit = (k[j] >> shift) & 255u;
                const int rank = __popc(peers[j] & lt);
                int pos = 0;
                if (valid) pos = sh.wcount[warp][digit] + rank;
                __syncwarp();
                if (valid && rank == 0) sh.wcount[warp][digit] += __popc(peers[j]);
                if (valid) { ko[pos] = k[j]; vo[pos] = vi[lo + 32 * j + lane]; }
                __syncwarp();
            }
        }
        flip ^= 1;
        __syncthreads();
    }
    return flip;
}

// The node list simulation. Node arrays live in shared memory (cap entries each).
struct QtNodes {
    int* lo[2]; int* hi[2]; int* dep[2];  // double-buffered list
    int* b1; int* b2; int* b3;            // inner child boundaries of expandable nodes
    int* cc;                              // non-empty children (0 for non-expandable)
    int* flag; int* scan;                 // scratch
    int* P;                               // processing order (node indices)
    int* inc;                             // per P-slot: children counts, then prefix sums
};

// core: candidates packed (x | y<<12 | score<<24) in cand[0..n); scratch = 4 arrays of n words.
// Writes the survivors (packed) in list order to sel[] and their number to *count_out.
// (no __restrict__ / read-only-cache hints: cand and scratch are written earlier in the same kernel)
// layout_cap >= sel_cap sizes the node arrays (the key area starts behind 14 * layout_cap ints of nodemem);
// keys_ready: the caller has already written path keys and candidate indices where qt_sort_input() points.
struct QtSortInput { uint32_t* keys; uint32_t* vals; bool small; };
__device__ __forceinline__ QtSortInput qt_sort_input(int n, uint32_t* scratch, int* nodemem, int layout_cap) {
    QtSortInput in;
    in.small = n <= kQtSmemKeys / 2;   // both buffer pairs of the sort are quarters of the shared-memory key area
    uint32_t* const sarea = reinterpret_cast<uint32_t*>(nodemem + 14 * layout_cap);
    in.keys = in.small ? sarea : scratch;
    in.vals = in.small ? sarea + kQtSmemKeys / 2 : scratch + n;
    return in;
}

__device__ void quadtree_select(const uint32_t* cand, int n, int N, int nRoots, float rootW, int H, int depth,
                                uint32_t* scratch, uint32_t* sel, int sel_cap, int* count_out,
                                QtShared& sh, int* nodemem, int layout_cap, bool keys_ready) {
    if (n <= 0 || nRoots <= 0) {
        if (threadIdx.x == 0) *count_out = 0;
        return;
    }
    uint32_t *kA = scratch, *vA = scratch + n, *kB = scratch + 2 * (size_t)n, *vB = scratch + 3 * (size_t)n;
    QT_MARK(1);
    const QtSortInput in = qt_sort_input(n, scratch, nodemem, layout_cap);
    const bool small = in.small;
    uint32_t* const sarea = reinterpret_cast<uint32_t*>(nodemem + 14 * layout_cap);
    if (small) { kA = in.keys; vA = in.vals; kB = sarea + kQtSmemKeys; vB = sarea + kQtSmemKeys + kQtSmemKeys / 2; }
    if (!keys_ready) {
        for (int i = threadIdx.x; i < n; i += kQtThreads) {
            const uint32_t p = cand[i];
            kA[i] = qt_path_key(p & 0xfff, (p >> 12) & 0xfff, H, rootW, depth);
            vA[i] = i;
        }
    }
    __syncthreads();
    int root_bits = 0;
    while ((1 << root_bits) < nRoots) ++root_bits;
    // Fast path (the usual case): sorted keys and a per-element selection word live in shared memory, so
    // the binary searches of the list simulation and the per-node maximum never leave the SM. The second
    // buffer pair of the sort IS that shared memory: with the usual three 8-bit passes the sorted set ends
    // there (global -> shared -> global -> shared) and only every other pass touches L2.
    const bool fast = n <= kQtSmemKeys;
    uint32_t* skey = sarea;
    uint32_t* comb = skey + kQtSmemKeys;  // response << 24 | (0xffffff - candidate index): max = best, earliest on ties
    if (fast && !small) { kB = skey; vB = comb; }
    QT_MARK(2);
    const int flip = small ? block_radix_sort_small(kA, vA, kB, vB, n, 2 * depth + root_bits, sh)
                           : block_radix_sort(kA, vA, kB, vB, n, 2 * depth + root_bits, sh);
    QT_MARK(3);
    const uint32_t* sk = flip ? kB : kA;
    const uint32_t* sv = flip ? vB : vA;
    if (small) {
        // sorted keys -> skey[0, n), selection words -> comb[0, n). Either buffer pair overlaps one of the two targets, but
        // element i of a source only ever shares its address with element i of a target, and each thread reads both of its
        // sources before it writes: no barrier needed in between.
        for (int i = threadIdx.x; i < n; i += kQtThreads) {
            const uint32_t kk = sk[i], c = sv[i];
            skey[i] = kk;
            comb[i] = (cand[c] >> 24) << 24 | (0xffffffu - c);
        }
        __syncthreads();
        sk = skey;
    } else if (fast) {
        if (flip) {  // already in shared memory: turn the candidate indices into selection words in place
            for (int i = threadIdx.x; i < n; i += kQtThreads) {
                const uint32_t c = comb[i];
                comb[i] = (cand[c] >> 24) << 24 | (0xffffffu - c);
            }
        } else {
            for (int i = threadIdx.x; i < n; i += kQtThreads) {
                const uint32_t c = sv[i];
                skey[i] = sk[i];
                comb[i] = (cand[c] >> 24) << 24 | (0xffffffu - c);
            }
        }
        __syncthreads();
        sk = skey;
    }

    const int cap = layout_cap;
    QtNodes q;
    {
        int* p = nodemem;
        for (int s = 0; s < 2; ++s) { q.lo[s] = p; p += cap; q.hi[s] = p; p += cap; q.dep[s] = p; p += cap; }
        q.b1 = p; p += cap; q.b2 = p; p += cap; q.b3 = p; p += cap; q.cc = p; p += cap;
        q.flag = p; p += cap; q.scan = p; p += cap; q.P = p; p += cap; q.inc = p; p += cap;
    }
    auto lower_bound_key = [&](uint32_t key) {
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sk[mid] >= key) hi = mid; else lo = mid + 1; }
        return lo;
    };

    // roots (539-585): non-empty ones, in order
    int cur = 0;
    int count;
    if (nRoots <= 32) {  // the usual 1 - 2 roots: one warp, one vote, one block barrier
        if (threadIdx.x < 32) {
            const int r = threadIdx.x;
            int lo = 0, hi = 0;
            if (r < nRoots) {
                lo = lower_bound_key((uint32_t)r << (2 * depth));
                hi = r + 1 == nRoots ? n : lower_bound_key((uint32_t)(r + 1) << (2 * depth));
            }
            const uint32_t nonempty = __ballot_sync(0xffffffffu, hi > lo);
            if (hi > lo) { const int p = __popc(nonempty & ((1u << r) - 1u)); q.lo[0][p] = lo; q.hi[0][p] = hi; q.dep[0][p] = 0; }
            if (r == 0) sh.bcast[0] = __popc(nonempty);
        }
        __syncthreads();
        count = sh.bcast[0];
    } else {
        for (int r = threadIdx.x; r < nRoots; r += kQtThreads) {
            const int lo = lower_bound_key((uint32_t)r << (2 * depth));
            const int hi = r + 1 == nRoots ? n : lower_bound_key((uint32_t)(r + 1) << (2 * depth));
            q.b1[r] = lo; q.b2[r] = hi; q.flag[r] = hi > lo;
        }
        __syncthreads();
        count = block_exclusive_scan(q.flag, q.scan, nRoots, sh);
        for (int r = threadIdx.x; r < nRoots; r += kQtThreads)
            if (q.flag[r]) { const int p = q.scan[r]; q.lo[0][p] = q.b1[r]; q.hi[0][p] = q.b2[r]; q.dep[0][p] = 0; }
        __syncthreads();
    }

    bool final_phase = false;
    QT_MARK(4);
    for (;;) {
        const int before = count;
        int *lo = q.lo[cur], *hi = q.hi[cur], *dep = q.dep[cur];
        int *nlo = q.lo[cur ^ 1], *nhi = q.hi[cur ^ 1], *ndep = q.dep[cur ^ 1];
        // children of every expandable node: one thread per (node, inner boundary)
        for (int it = threadIdx.x; it < 3 * before; it += kQtThreads) {
            const int i = it / 3, t = it - 3 * i + 1;
            const int l = lo[i], h = hi[i], d = dep[i];
            if (h - l >= 2 && d < depth) {
                const int sft = 2 * (depth - d - 1);
                int a = l, b = h;
                while (a < b) { const int mid = (a + b) >> 1; if (((sk[mid] >> sft) & 3u) >= (uint32_t)t) b = mid; else a = mid + 1; }
                (t == 1 ? q.b1 : t == 2 ? q.b2 : q.b3)[i] = a;
            }
        }
        __syncthreads();
        if (!final_phase && before <= kQtThreads) {
            // A full pass over a list that fits one node per thread (the usual case): every expandable node is processed
            // in list order, so the place of its children (pushed to the front, later nodes first) and of every untouched
            // node follows from ONE scan over (children | expandable | expandable children) packed in a word - three block
            // barriers per pass instead of ~25 (the general form below also serves the final, size-ordered phase).
            const int i = threadIdx.x, lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
            int l = 0, h = 0, d = 0, c = 0, ex = 0, me = 0, bb[5] = {0, 0, 0, 0, 0};
            if (i < before) {
                l = lo[i]; h = hi[i]; d = dep[i];
                if (h - l >= 2 && d < depth) {
                    bb[0] = l; bb[1] = q.b1[i]; bb[2] = q.b2[i]; bb[3] = q.b3[i]; bb[4] = h;
                    ex = 1;
#pragma unroll
                    for (int t = 0; t < 4; ++t) { c += bb[t + 1] > bb[t]; me += bb[t + 1] - bb[t] >= 2; }
                }
            }
            const int v = c | ex << 11 | me << 20;   // sums stay below 2^11, 2^9, 2^11
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane_ >= o) inc += t; }
            if (lane_ == 31) sh.warp_tmp[warp_] = inc;
            __syncthreads();
            int off = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) { const int t = sh.warp_tmp[w]; total += t; off += w < warp_ ? t : 0; }
            const int excl = off + inc - v;
            const int children = total & 0x7ff, m = (total >> 11) & 0x1ff, n_expand = total >> 20;
            if (i < before) {
                if (ex) {
                    int pos = children - (excl & 0x7ff) - c;
#pragma unroll
                    for (int t = 3; t >= 0; --t)
                        if (bb[t + 1] > bb[t]) { nlo[pos] = bb[t]; nhi[pos] = bb[t + 1]; ndep[pos] = d + 1; ++pos; }
                } else {
                    const int p = children + i - ((excl >> 11) & 0x1ff);
                    nlo[p] = l; nhi[p] = h; ndep[p] = d;
                }
            }
            __syncthreads();  // the new list is complete; warp_tmp may be reused
            count = children + (before - m);
            cur ^= 1;
            QT_MARK(10000 + before);
            if (count >= N || count == before) break;
            if (count + 3 * n_expand > N) final_phase = true;
            continue;
        }
        if (final_phase && before <= kQtThreads) {
            // The size-ordered phase (676-737) on a list that fits one node per thread: slot of an expandable node = its rank by
            // (size descending, list position ascending); the children counts are scattered to the slots and scanned there; the
            // number of slots that get processed is the first at which the list reaches N (730-731); children of processed
            // nodes go to the front (later slots first), everything else follows in list order. 7 block barriers instead of ~25.
            const int i = threadIdx.x, lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
            int l = 0, h = 0, d = 0, c = 0, ex = 0, bb[5] = {0, 0, 0, 0, 0};
            if (i < before) {
                l = lo[i]; h = hi[i]; d = dep[i];
                if (h - l >= 2 && d < depth) {
                    bb[0] = l; bb[1] = q.b1[i]; bb[2] = q.b2[i]; bb[3] = q.b3[i]; bb[4] = h;
                    ex = 1;
#pragma unroll
                    for (int t = 0; t < 4; ++t) c += bb[t + 1] > bb[t];
                }
            }
            if (i < before) q.flag[i] = ex ? h - l : 0;                               // sizes of the expandable nodes (>= 2), 0 otherwise
            if (threadIdx.x == 0) sh.bcast[1] = 0x7fffffff;
            const int m = __syncthreads_count(ex);                                    // also publishes the sizes
            int rank = -1;
            if (ex) {
                const int mine = h - l;
                rank = 0;
                for (int j = 0; j < before; ++j) { const int sj = q.flag[j]; rank += (sj > mine) || (sj == mine && j < i); }
                q.inc[rank] = c;                                                      // children count of slot `rank`
            }
            __syncthreads();
            // exclusive scan of the children counts over the slots (thread k = slot k)
            const int v = i < m ? q.inc[i] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane_ >= o) inc += t; }
            if (lane_ == 31) sh.warp_tmp[warp_] = inc;
            __syncthreads();
            int off = 0, all_children = 0;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) { const int t = sh.warp_tmp[w]; all_children += t; off += w < warp_ ? t : 0; }
            const int before_k = off + inc - v;                                       // children of the slots before slot i
            if (i < m) {
                q.scan[i] = before_k;
                if (before + before_k + v - (i + 1) >= N) atomicMin(&sh.bcast[1], i + 1);   // list size after processing slots 0..i
            }
            __syncthreads();
            const int processed = min(sh.bcast[1], m);
            const int children = processed < m ? q.scan[processed] : all_children;
            // untouched nodes keep their order behind the children: their rank among themselves by one more scan
            const int untouched = i < before && !(ex && rank < processed);
            int uinc = untouched;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, uinc, o); if (lane_ >= o) uinc += t; }
            __syncthreads();                                                          // warp_tmp is free again
            if (lane_ == 31) sh.warp_tmp[warp_] = uinc;
            __syncthreads();
            int uoff = 0;
#pragma unroll
            for (int w = 0; w < kQtWarps; ++w) uoff += w < warp_ ? sh.warp_tmp[w] : 0;
            if (i < before) {
                if (!untouched) {
                    int pos = children - q.scan[rank] - c;
#pragma unroll
                    for (int t = 3; t >= 0; --t)
                        if (bb[t + 1] > bb[t]) { nlo[pos] = bb[t]; nhi[pos] = bb[t + 1]; ndep[pos] = d + 1; ++pos; }
                } else {
                    const int p = children + uoff + uinc - 1;
                    nlo[p] = l; nhi[p] = h; ndep[p] = d;
                }
            }
            __syncthreads();  // the new list is complete; warp_tmp, flag, inc, scan may be reused
            count = children + (before - processed);
            cur ^= 1;
            QT_MARK(100000 + before);
            if (count >= N || count == before) break;
            continue;
        }
        for (int i = threadIdx.x; i < before; i += kQtThreads) {
            const int l = lo[i], h = hi[i];
            int c = 0;
            if (h - l >= 2 && dep[i] < depth) c = (q.b1[i] > l) + (q.b2[i] > q.b1[i]) + (q.b3[i] > q.b2[i]) + (h > q.b3[i]);
            q.cc[i] = c;
            q.flag[i] = c > 0;
        }
        __syncthreads();
        const int m = block_exclusive_scan(q.flag, q.scan, before, sh);
        // processing order P
        if (!final_phase) {
            for (int i = threadIdx.x; i < before; i += kQtThreads) if (q.flag[i]) q.P[q.scan[i]] = i;
        } else {
            // compact list-ordered expandable nodes into inc[] (temporarily), then rank by
            // (size desc, list position asc)
            for (int i = threadIdx.x; i < before; i += kQtThreads) if (q.flag[i]) q.inc[q.scan[i]] = i;
            __syncthreads();
            for (int a = threadIdx.x; a < m; a += kQtThreads) {
                const int na = q.inc[a], sa = hi[na] - lo[na];
                int rank = 0;
                for (int b = 0; b < m; ++b) {
                    const int nb = q.inc[b], sb = hi[nb] - lo[nb];
                    rank += (sb > sa) || (sb == sa && b < a);
                }
                q.P[rank] = na;
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < m; k += kQtThreads) q.inc[k] = q.cc[q.P[k]];
        __syncthreads();
        const int all_children = block_exclusive_scan(q.inc, q.inc, m, sh);  // inc[k] = children before slot k
        // how many slots get processed (730-731)
        int processed = m;
        if (final_phase) {
            if (threadIdx.x == 0) sh.bcast[1] = m;
            __syncthreads();
            for (int k = threadIdx.x; k < m; k += kQtThreads) {
                // list size after processing slots 0..k
                const int after_k = before + q.inc[k] + q.cc[q.P[k]] - (k + 1);
                if (after_k >= N) atomicMin(&sh.bcast[1], k + 1);
            }
            __syncthreads();
            processed = sh.bcast[1];
        }
        const int children = processed < m ? q.inc[processed] : all_children;
        __syncthreads();
        // new list: children of slot processed-1 (q=3..0), ..., of slot 0; then untouched nodes in order
        for (int i = threadIdx.x; i < before; i += kQtThreads) q.flag[i] = 1;  // 1 = untouched
        if (threadIdx.x == 0) sh.bcast[2] = 0;
        __syncthreads();
        int my_expand = 0;
        for (int k = threadIdx.x; k < processed; k += kQtThreads) {
            const int i = q.P[k];
            q.flag[i] = 0;
            const int bnd[5] = {lo[i], q.b1[i], q.b2[i], q.b3[i], hi[i]};
            int pos = children - q.inc[k] - q.cc[i];
            const int d = dep[i] + 1;
#pragma unroll
            for (int t = 3; t >= 0; --t)
                if (bnd[t + 1] > bnd[t]) {
                    nlo[pos] = bnd[t]; nhi[pos] = bnd[t + 1]; ndep[pos] = d; ++pos;
                    my_expand += (bnd[t + 1] - bnd[t]) >= 2;
                }
        }
        if (my_expand) atomicAdd(&sh.bcast[2], my_expand);
        __syncthreads();
        block_exclusive_scan(q.flag, q.scan, before, sh);
        for (int i = threadIdx.x; i < before; i += kQtThreads)
            if (q.flag[i]) { const int p = children + q.scan[i]; nlo[p] = lo[i]; nhi[p] = hi[i]; ndep[p] = dep[i]; }
        __syncthreads();
        const int n_expand = sh.bcast[2];
        count = children + (before - processed);
        cur ^= 1;
        QT_MARK(final_phase ? 100000 + before : 10000 + before);
        if (count >= N || count == before) break;
        if (!final_phase && count + 3 * n_expand > N) final_phase = true;
        __syncthreads();
    }

    // best candidate per node: max response, earliest candidate on ties
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* lo = q.lo[cur]; const int* hi = q.hi[cur];
    if (fast) {  // one thread per node over the shared-memory selection words
        for (int p = threadIdx.x; p < count; p += kQtThreads) {
            uint32_t best = 0;
            for (int i = lo[p]; i < hi[p]; ++i) best = max(best, comb[i]);
            if (p < sel_cap) sel[p] = cand[0xffffffu - (best & 0xffffffu)];
        }
    } else {     // one warp per node over the global arrays
        for (int p = warp; p < count; p += kQtWarps) {
            uint32_t best = 0;
            for (int i = lo[p] + lane; i < hi[p]; i += 32) {
                const uint32_t c = sv[i];
                best = max(best, (cand[c] >> 24) << 24 | (0xffffffu - c));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
            if (lane == 0 && p < sel_cap) sel[p] = cand[0xffffffu - (best & 0xffffffu)];
        }
    }
    QT_MARK(5);
    if (threadIdx.x == 0) *count_out = min(count, sel_cap);
#ifdef ORB_QT_PROFILE
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        printf("qt n=%d N=%d depth=%d:", n, N, depth);
        for (int i = 1; i < g_qt_nmarks; ++i) printf(" [%d]%lld", g_qt_tags[i], g_qt_marks[i] - g_qt_marks[i - 1]);
        printf("\n");
        g_qt_nmarks = 0;
    }
#endif
}

static size_t qt_smem_bytes(int sel_cap) { return (size_t)sel_cap * 14 * sizeof(int) + 2 * (size_t)kQtSmemKeys * sizeof(uint32_t); }

// grid (levels, frames)
__global__ void __launch_bounds__(kQtThreads, 4)
quadtree_kernel(const Geometry* __restrict__ g, const uint32_t* __restrict__ slots, const int* __restrict__ cell_counts,
                uint32_t* sortbuf, uint32_t* selected, int* sel_counts, int level0) {
    extern __shared__ __align__(16) int nodemem[];
    __shared__ QtShared sh;
    const int level = level0 + blockIdx.x, frame = blockIdx.y;
    const LevelGeom& L = g->lv[level];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QT_MARK(0);
    pdl_release_dependents();   // the description kernel may be set up (its tables filled) while the trees are built
    pdl_wait();                 // programmatic dependent of the FAST launch: its candidate lists are complete from here on
    uint32_t* fb = sortbuf + (size_t)frame * 5 * g->cand_words;
    uint32_t* cand = fb + L.cand_off;
    uint32_t* scratch = fb + g->cand_words + 4 * L.cand_off;
    const int* counts = cell_counts + (size_t)frame * g->ncells + L.cell_begin;
    const uint32_t* lslots = slots + (size_t)frame * g->slot_words + L.slot_off;

    // stitch the cells' lists in row-major cell order (789-829): offsets = scan of counts.
    // nodemem is free until quadtree_select: cell counts (one coalesced pass) and their offsets live there
    // (14 * layout_cap ints >= 2 * cell_count)
    const int layout_cap = max(L.sel_cap, (2 * L.cell_count + 13) / 14 + 1);
    int* offs = nodemem;
    int* cnt_s = nodemem + L.cell_count;
    for (int c = threadIdx.x; c < L.cell_count; c += kQtThreads) cnt_s[c] = counts[c];
    const int total = block_exclusive_scan(cnt_s, offs, L.cell_count, sh);
    __syncthreads();
    // 8 lanes per cell (a cell holds ~4 candidates, rarely more than 8), two cell groups per step: the slot loads of 8 cells per
    // warp are in flight together
    for (int c0 = warp * 4; c0 < L.cell_count; c0 += 2 * kQtWarps * 4) {
        const int ca = c0 + (lane >> 3), cb = ca + kQtWarps * 4, k0 = lane & 7;
        const int na = ca < L.cell_count ? cnt_s[ca] : 0, nb = cb < L.cell_count ? cnt_s[cb] : 0;
        const int oa = ca < L.cell_count ? offs[ca] : 0, ob = cb < L.cell_count ? offs[cb] : 0;
        const uint32_t* sa = lslots + (size_t)ca * L.slot_cap;
        const uint32_t* sb = lslots + (size_t)cb * L.slot_cap;
        const uint32_t pa = k0 < na ? sa[k0] : 0u, pb = k0 < nb ? sb[k0] : 0u;   // both loads before either store
        if (k0 < na) cand[oa + k0] = pa;
        if (k0 < nb) cand[ob + k0] = pb;
        for (int k = k0 + 8; k < na; k += 8) cand[oa + k] = sa[k];
        for (int k = k0 + 8; k < nb; k += 8) cand[ob + k] = sb[k];
    }
    __syncthreads();
    const int H = L.h - 2 * kMinBorder;
    quadtree_select(cand, total, L.quota, L.nRoots, L.rootW, H, L.key_depth, scratch,
                    selected + (size_t)frame * g->sel_words + L.sel_off, L.sel_cap,
                    sel_counts + (size_t)frame * g->nlevels + level, sh, nodemem, layout_cap, false);
}

__global__ void __launch_bounds__(kQtThreads)
quadtree_standalone_kernel(const uint32_t* cand, int n, int N, int nRoots, float rootW, int H, int depth,
                           uint32_t* scratch, uint32_t* sel, int sel_cap, int* count) {
    extern __shared__ __align__(16) int nodemem[];
    __shared__ QtShared sh;
    quadtree_select(cand, n, N, nRoots, rootW, H, depth, scratch, sel, sel_cap, count, sh, nodemem, sel_cap, false);
}

// opt in once per device to the large dynamic shared memory carve-out
constexpr size_t kQtMaxDynSmem = 160 * 1024;
static int qt_configure(size_t need) {
    static std::atomic<bool> done[64];  // idempotent per device; a racing second call only repeats the attribute set
    if (need > kQtMaxDynSmem) { set_error("nfeatures too large for the quadtree kernel's shared memory"); return ORB_EINVAL; }
    int dev = 0;
    ORB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return ORB_OK;
    ORB_CUDA_TRY(cudaFuncSetAttribute(quadtree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQtMaxDynSmem));
    ORB_CUDA_TRY(cudaFuncSetAttribute(quadtree_standalone_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQtMaxDynSmem));
    if (dev < 64) done[dev] = true;
    return ORB_OK;
}

// level < 0: all levels (grid.x = levels); otherwise that level only
int launch_quadtree(const Geometry& hg, const DeviceBuffers& db, int n, cudaStream_t st, int level) {
    int max_cap = 0;
    for (int l = 0; l < hg.nlevels; ++l) max_cap = max(max_cap, max(hg.lv[l].sel_cap, ceil_div(2 * hg.lv[l].cell_count, 14) + 1));
    const size_t smem = qt_smem_bytes(max_cap);
    int rc = qt_configure(smem);
    if (rc) return rc;
    // always a programmatic dependent of the FAST launch before it in the stream (the kernel waits before its first read)
    ORB_CUDA_TRY(launch_pdl(quadtree_kernel, dim3(level < 0 ? hg.nlevels : 1, n), dim3(kQtThreads), smem, st, db.geom, db.slots, db.cell_counts,
                            db.sortbuf, db.selected, db.sel_counts, level < 0 ? 0 : level));
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int launch_quadtree_standalone(const uint32_t* d_cand, int n, int N, int nRoots, float rootW, int H, int depth,
                               uint32_t* d_scratch4n, uint32_t* d_sel, int sel_cap, int* d_count, cudaStream_t st) {
    const size_t smem = qt_smem_bytes(sel_cap);
    int rc = qt_configure(smem);
    if (rc) return rc;
    quadtree_standalone_kernel<<<1, kQtThreads, smem, st>>>(d_cand, n, N, nRoots, rootW, H, depth, d_scratch4n, d_sel, sel_cap, d_count);
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orb
