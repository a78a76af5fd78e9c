// Extractor handle: geometry tables (the reference ctor, /root/reference/src/ORBextractor.cc:410-470,
// plus the per-level cell grid of ComputeKeyPointsOctTree 771-787), HBM buffers, kernel
// orchestration (operator(), 1043-1105) and the orbx_* C ABI.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "extract_kernels.cuh"

namespace orb {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
int tma_encode_u8_3d(CUtensorMap* out, const void* base, int w, int h, int frames, size_t pitch, size_t frame_stride, int bw, int bh) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::atomic<EncodeFn> cached{nullptr};  // resolved once; racing threads resolve the same pointer
    EncodeFn fn = cached.load();
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        ORB_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return ORB_ECUDA; }
        fn = (EncodeFn)p;
        cached.store(fn);
    }
    if (((uintptr_t)base & 15) || (pitch & 15) || (frame_stride & 15) || bw > 256 || bh > 256 || (bw & 15)) {
        set_error("tensor map: base/strides must be 16-byte aligned, box <= 256 (pitch %zu, stride %zu, box %dx%d)", pitch, frame_stride, bw, bh);
        return ORB_EINVAL;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return ORB_ECUDA; }
    return ORB_OK;
}

static const int kPatternInts[1024] = {
#include "orb_pattern_31.inc"
};

// cvRound / cvFloor / cvCeil of the ctor
static inline int round_half_even(float v) { return (int)nearbyintf(v); }
static inline int round_half_even_d(double v) { return (int)nearbyint(v); }
static inline int ifloor(double v) { int i = (int)v; return i - (i > v); }
static inline int iceil(double v) { int i = (int)v; return i + (i < v); }

static std::vector<LinTap> linear_taps(int src_n, int dst_n) {
    std::vector<LinTap> t(dst_n);
    const double scale = (double)src_n / dst_n;
    for (int d = 0; d < dst_n; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = ifloor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
        t[d].ofs = s;
        t[d].c0 = (short)round_half_even((1.f - f) * 2048.f);
        t[d].c1 = (short)round_half_even(f * 2048.f);
    }
    return t;
}

}  // namespace orb

using namespace orb;

struct orbx_extractor {
    orbx_config cfg;
    int device = 0, width = 0, height = 0, max_batch = 0;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> quota;
    Geometry hg;  // host copy
    DeviceBuffers db{};
    void* d_geom = nullptr; void* d_cells = nullptr; void* d_taps = nullptr; void* d_tiles = nullptr; void* d_pattern = nullptr;
    uint8_t* d_input = nullptr;  // staging for host frames: [max_batch][height][pitch0]
    float* d_stereo = nullptr;   // outputs of the host-buffer stereo call (uRight, depth, sad, kept), allocated on first use
    uint8_t* h_pyr = nullptr;    // pinned staging of orbx_pyramid_levels (all levels of one frame, tightly packed), on first use
    size_t in_pitch = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev_pyr = nullptr, ev_blur = nullptr;
    uint8_t* d_results = nullptr;                      // counts | kps | desc (db.counts / db.kps / db.desc point into it)
    size_t results_kps_off() const { return align_up((size_t)max_batch * 4, 256); }
    size_t results_bytes() const { return results_kps_off() + (size_t)max_batch * hg.out_cap * (sizeof(orbx_keypoint) + 32); }
    uint8_t* h_out = nullptr;                          // pinned staging of the results for small synchronous host calls
    size_t h_out_bytes = 0;
    cudaStream_t lvl_stream[kMaxLevels] = {};          // small batches: one branch per pyramid level (launch_sequence)
    cudaEvent_t ev_lvl_ready[kMaxLevels] = {}, ev_lvl_done[kMaxLevels] = {};
    FrameSet last{};  // frames of the last call (for mvImagePyramid level 0)
    int last_n = 0;
    // TMA tensor maps per level: source windows of the resize, blur tiles, FAST cell tiles.
    // Levels >= 1 are encoded once; level 0 follows the frames of the call (cached on the FrameSet).
    TmaMaps maps_resize, maps_blur, maps_fast;
    FrameSet maps_l0{};
    // CUDA graphs of the kernel sequence for small batches (the per-frame operator() path is launch bound: 11 launches +
    // fork/join). A small LRU keyed on (frames, n) - a caller alternating two input buffers keeps both graphs; state
    // 0 = free, 1 = ran eagerly once, 2 = captured. Capture failures are tolerated a few times (e.g. the caller's stream
    // is itself being captured) before graphs are given up for the handle.
    struct GraphEntry { FrameSet fs{}; int n = 0, state = 0, kernels = 0; cudaGraphExec_t exec = nullptr; unsigned long long stamp = 0; };
    static constexpr int kGraphEntries = 4;
    GraphEntry graphs[kGraphEntries];
    unsigned long long graph_clock = 0;
    int graph_failures = 0;
    bool use_graphs = true;
    // optional per-stage device timing (bench roofline): events around each stage of the pipeline
    bool stage_timing = false;
    cudaEvent_t ev_stage[ORBX_NUM_STAGES + 1] = {};
    cudaEvent_t ev_blur_begin = nullptr, ev_blur_end = nullptr;
};

namespace {

int build_geometry(orbx_extractor* h) {
    const orbx_config& c = h->cfg;
    const int nl = c.nlevels;
    // ---- ctor tables (410-446): float chain with a double scaleFactor member ----------------
    const double sf = (double)c.scale_factor;
    h->scale.assign(nl, 1.f); h->sigma2.assign(nl, 1.f); h->inv_scale.resize(nl); h->inv_sigma2.resize(nl);
    for (int i = 1; i < nl; ++i) { h->scale[i] = (float)(h->scale[i - 1] * sf); h->sigma2[i] = h->scale[i] * h->scale[i]; }
    for (int i = 0; i < nl; ++i) { h->inv_scale[i] = 1.0f / h->scale[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    h->quota.resize(nl);
    {
        const float factor = (float)(1.0f / sf);
        float per = c.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) { h->quota[l] = round_half_even(per); sum += h->quota[l]; per *= factor; }
        h->quota[nl - 1] = std::max(c.nfeatures - sum, 0);
    }
    Geometry& g = h->hg;
    memset(&g, 0, sizeof(g));
    g.nlevels = nl; g.iniTh = c.ini_th_fast; g.minTh = c.min_th_fast;
    {   // umax (452-469)
        int v, v0;
        const int vmax = ifloor(kHalfPatch * sqrt(2.f) / 2 + 1), vmin = iceil(kHalfPatch * sqrt(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (v = 0; v <= vmax; ++v) g.umax[v] = round_half_even_d(sqrt(hp2 - v * v));
        for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (g.umax[v0] == g.umax[v0 + 1]) ++v0;
            g.umax[v] = v0;
            ++v0;
        }
    }
    std::vector<CellDesc> cells;
    std::vector<LinTap> taps;
    std::vector<BlurTile> tiles;
    size_t pyr = 0, blur = 0, slot = 0;
    int sel = 0;
    for (int l = 0; l < nl; ++l) {
        LevelGeom& L = g.lv[l];
        L.w = round_half_even((float)h->width * h->inv_scale[l]);   // 1111-1112
        L.h = round_half_even((float)h->height * h->inv_scale[l]);
        L.pitch = (int)align_up((size_t)L.w, 128);
        L.scale = h->scale[l];
        L.patch_size = (int)(31 * h->scale[l]);  // 837
        L.quota = h->quota[l];
        const int minB = kMinBorder, maxBX = L.w - kEdgeThreshold + 3, maxBY = L.h - kEdgeThreshold + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        L.nCols = (int)(width / 30.f); L.nRows = (int)(height / 30.f);  // 784-785
        if (L.nCols < 1 || L.nRows < 1) { set_error("pyramid level %d (%dx%d) is smaller than one 30 px cell", l, L.w, L.h); return ORB_EINVAL; }
        L.wCell = (int)ceilf(width / L.nCols); L.hCell = (int)ceilf(height / L.nRows);
        L.nRoots = (int)roundf(width / (float)(maxBY - minB));  // 543
        if (L.nRoots < 1) { set_error("level %d: aspect ratio below 0.5 is not supported by the reference quadtree", l); return ORB_EINVAL; }
        L.rootW = width / L.nRoots;  // 545
        int root_bits = 0;
        while ((1 << root_bits) < L.nRoots) ++root_bits;
        const int span = std::max((int)ceilf(L.rootW) + 1, maxBY - minB);
        int depth = 1;
        while ((1 << depth) < span) ++depth;
        L.key_depth = std::min(depth + 1, kQtMaxDepth);
        if (span > 4096 || 2 * L.key_depth + root_bits > 31) { set_error("level %d too large for the 32-bit quadtree key", l); return ORB_EINVAL; }
        L.slot_cap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);  // one NMS survivor per 2x2 block at most
        L.cell_begin = (int)cells.size();
        int ord = 0;
        for (int i = 0; i < L.nRows; ++i) {  // 789-806
            const int y0 = minB + i * L.hCell;
            int y1 = y0 + L.hCell + 6;
            if (y0 >= maxBY - 3) continue;
            if (y1 > maxBY) y1 = maxBY;
            for (int j = 0; j < L.nCols; ++j) {
                const int x0 = minB + j * L.wCell;
                int x1 = x0 + L.wCell + 6;
                if (x0 >= maxBX - 6) continue;
                if (x1 > maxBX) x1 = maxBX;
                CellDesc cd{};
                cd.level = (int16_t)l; cd.x0 = (int16_t)x0; cd.y0 = (int16_t)y0; cd.tw = (int16_t)(x1 - x0); cd.th = (int16_t)(y1 - y0);
                cd.offx = (int16_t)(j * L.wCell); cd.offy = (int16_t)(i * L.hCell); cd.ordinal = ord++;
                {   // FAST phase 1 walks the interior columns in aligned 4-byte groups of the tile row (tile byte 0 = x0 & ~15)
                    const int bs = (x0 & 15) + 3, be = bs + (x1 - x0) - 6;
                    const int G = be > bs ? ((be - 1) >> 2) - (bs >> 2) + 1 : 1;
                    cd.ginv = (65536u + G - 1) / G;
                }
                cells.push_back(cd);
                g.max_tw = std::max(g.max_tw, (int)cd.tw); g.max_th = std::max(g.max_th, (int)cd.th);
            }
        }
        L.cell_count = ord;
        L.cand_cap = ord * L.slot_cap;
        L.slot_off = slot; L.cand_off = slot;
        slot += (size_t)L.cand_cap;
        L.sel_cap = std::max(L.quota, 4 * L.nRoots) + 3;
        L.sel_off = sel; sel += L.sel_cap;
        L.blur_off = blur; blur += align_up((size_t)L.pitch * L.h, 256);
        if (l > 0) {
            L.img_off = pyr; pyr += align_up((size_t)L.pitch * L.h, 256);
            std::vector<LinTap> tx = linear_taps(g.lv[l - 1].w, L.w), ty = linear_taps(g.lv[l - 1].h, L.h);
            L.tab_x_off = (int)taps.size(); taps.insert(taps.end(), tx.begin(), tx.end());
            L.tab_y_off = (int)taps.size(); taps.insert(taps.end(), ty.begin(), ty.end());
        }
        for (int ty = 0; ty < ceil_div(L.h, 58); ++ty)  // kBlurTH x kBlurTW output tiles (pyramid.cu)
            for (int tx = 0; tx < ceil_div(L.w, 128); ++tx) tiles.push_back(BlurTile{(int16_t)l, (int16_t)tx, (int16_t)ty, 0});
    }
    g.ncells = (int)cells.size(); g.ntiles = (int)tiles.size();
    // the innermost TMA coordinate is 16-byte granular: boxes carry up to 15 lead-in bytes
    g.fast_bw = (int)align_up((size_t)g.max_tw + 15, 16);
    g.rs_bw = (int)align_up((size_t)ceil(128.0 * sf) + 2 + 15 + 12, 16);  // + 3-word read window of the last column quad
    g.rs_bh = (int)ceil(32.0 * sf) + 3;  // kRsTH output rows (pyramid.cu)
    if (g.fast_bw > 256 || g.max_th > 256 || g.rs_bw > 256) { set_error("scale factor / cell size too large for the TMA tile boxes"); return ORB_EINVAL; }
    g.pyr_bytes = std::max<size_t>(pyr, 256); g.blur_bytes = blur; g.slot_words = slot; g.cand_words = slot;
    g.sel_words = sel; g.out_cap = sel;
    if (taps.empty()) taps.push_back(LinTap{0, 0, 0});

    const int B = h->max_batch;
    auto dalloc = [&](void** p, size_t bytes) -> int { ORB_CUDA_TRY(cudaMalloc(p, std::max<size_t>(bytes, 256))); return ORB_OK; };
    int rc;
    if ((rc = dalloc(&h->d_geom, sizeof(Geometry))) || (rc = dalloc(&h->d_cells, cells.size() * sizeof(CellDesc))) ||
        (rc = dalloc(&h->d_taps, taps.size() * sizeof(LinTap))) || (rc = dalloc(&h->d_tiles, tiles.size() * sizeof(BlurTile))) ||
        (rc = dalloc(&h->d_pattern, (kPatternWords + kIcTableWords) * 4)))
        return rc;
    // sampling pattern for the describe kernel: the four int8 coordinates (x0, y0, x1, y1; |value| <= 13) of test t of
    // descriptor byte b in one word at [t * 32 + b] (1 KB of shared memory instead of 4 KB of floats)
    uint32_t patT[kPatternWords];
    for (int b = 0; b < 32; ++b)
        for (int t = 0; t < 8; ++t) {
            uint32_t w = 0;
            for (int c4 = 0; c4 < 4; ++c4) w |= (uint32_t)(uint8_t)(int8_t)kPatternInts[4 * (8 * b + t) + c4] << (8 * c4);
            patT[t * 32 + b] = w;
        }
    ORB_CUDA_TRY(cudaMemcpy(h->d_geom, &g, sizeof(Geometry), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(h->d_cells, cells.data(), cells.size() * sizeof(CellDesc), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(h->d_taps, taps.data(), taps.size() * sizeof(LinTap), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(h->d_tiles, tiles.data(), tiles.size() * sizeof(BlurTile), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(h->d_pattern, patT, sizeof(patT), cudaMemcpyHostToDevice));
    {   // IC_Angle coefficient words (describe.cu): for each alignment phase p of the patch's left edge, item i = row * 9 + word:
        // u + 16 (1 .. 31) of the word's 4 bytes, 0 outside the circular patch. v is the same for a whole row, so
        // m01 = sum_rows v * sum_inside I needs no table of its own: "inside" is "byte != 0".
        std::vector<uint32_t> tab(kIcTableWords, 0u);
        for (int p = 0; p < 4; ++p)
            for (int r = 0; r < 2 * kHalfPatch + 1; ++r)
                for (int j = 0; j < kIcWordsPerRow; ++j) {
                    uint32_t cu = 0;
                    const int v = r - kHalfPatch;
                    for (int k = 0; k < 4; ++k) {
                        const int u = 4 * j + k - kHalfPatch - p;
                        if (u < -kHalfPatch || u > kHalfPatch || std::abs(u) > g.umax[std::abs(v)]) continue;
                        cu |= (uint32_t)(u + 16) << (8 * k);
                    }
                    tab[(size_t)p * kIcItems + r * kIcWordsPerRow + j] = cu;
                }
        ORB_CUDA_TRY(cudaMemcpy((uint8_t*)h->d_pattern + sizeof(patT), tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    }
    DeviceBuffers& db = h->db;
    db.geom = (const Geometry*)h->d_geom; db.cells = (const CellDesc*)h->d_cells; db.taps = (const LinTap*)h->d_taps;
    db.tiles = (const BlurTile*)h->d_tiles; db.pattern = (const uint32_t*)h->d_pattern;
    db.ic_table = (const uint32_t*)h->d_pattern + kPatternWords;
    h->in_pitch = align_up((size_t)h->width, 128);
    if ((rc = dalloc((void**)&db.pyr, (size_t)B * g.pyr_bytes)) || (rc = dalloc((void**)&db.blur, (size_t)B * g.blur_bytes)) ||
        (rc = dalloc((void**)&db.slots, (size_t)B * g.slot_words * 4)) || (rc = dalloc((void**)&db.cell_counts, (size_t)B * g.ncells * 4)) ||
        (rc = dalloc((void**)&db.sortbuf, (size_t)B * 5 * g.cand_words * 4)) || (rc = dalloc((void**)&db.selected, (size_t)B * g.sel_words * 4)) ||
        (rc = dalloc((void**)&db.sel_counts, (size_t)B * nl * 4)) || (rc = dalloc((void**)&h->d_results, h->results_bytes())) ||
        (rc = dalloc((void**)&h->d_input, (size_t)B * h->in_pitch * h->height)))
        return rc;
    // counts | keypoints | descriptors of all frames in ONE allocation: a call that uses the whole handle (always the case
    // for max_batch = 1, the facade) fetches its results with a single copy
    db.counts = (int*)h->d_results;
    db.kps = (orbx_keypoint*)(h->d_results + h->results_kps_off());
    db.desc = h->d_results + h->results_kps_off() + (size_t)B * g.out_cap * sizeof(orbx_keypoint);
    ORB_CUDA_TRY(cudaMemset(db.counts, 0, (size_t)B * 4));
    for (int l = 1; l < nl; ++l) {
        const LevelGeom& L = g.lv[l];
        const uint8_t* base = db.pyr + L.img_off;
        if ((rc = tma_encode_u8_3d(&h->maps_resize.m[l], base, L.w, L.h, B, L.pitch, g.pyr_bytes, g.rs_bw, g.rs_bh)) ||
            (rc = tma_encode_u8_3d(&h->maps_blur.m[l], base, L.w, L.h, B, L.pitch, g.pyr_bytes, 160, 64)) ||
            (rc = tma_encode_u8_3d(&h->maps_fast.m[l], base, L.w, L.h, B, L.pitch, g.pyr_bytes, g.fast_bw, g.max_th)))
            return rc;
    }
    ORB_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    ORB_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    ORB_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_pyr, cudaEventDisableTiming));
    ORB_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_blur, cudaEventDisableTiming));
    for (int l = 0; l < nl; ++l) {
        ORB_CUDA_TRY(cudaStreamCreateWithFlags(&h->lvl_stream[l], cudaStreamNonBlocking));
        ORB_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_lvl_ready[l], cudaEventDisableTiming));
        ORB_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_lvl_done[l], cudaEventDisableTiming));
    }
    return ORB_OK;
}

// the kernel sequence of operator(): resize chain, then blur on the side stream next to FAST + quadtree, joined
// before orientation + description
// Up to this many frames per call the GPU is far from full (one frame: 1300 FAST warps, 8 quadtree blocks) and the call's
// latency is the dependency chain resize 1..7 -> FAST -> quadtree -> describe (37 + 15 + 41 + 16 us at batch 1). Level l's
// detection and distribution only need pyramid level l, so they run as one branch per level next to the resize chain:
// the critical path becomes FAST + quadtree of level 0 (the input image itself) -> describe.
constexpr int kLevelParallelMaxBatch = 2;

static int launch_sequence_level_parallel(orbx_extractor* h, const FrameSet& fs, int n, cudaStream_t st) {
    const Geometry& g = h->hg;
    int rc;
    for (int l = 0; l < g.nlevels; ++l) {
        if (l > 0 && (rc = launch_resize_level(g, h->db, h->maps_resize, l, n, st, l > 1))) return rc;
        cudaStream_t s = h->lvl_stream[l];
        ORB_CUDA_TRY(cudaEventRecord(h->ev_lvl_ready[l], st));
        ORB_CUDA_TRY(cudaStreamWaitEvent(s, h->ev_lvl_ready[l], 0));
        if ((rc = launch_fast(g, h->db, h->maps_fast, n, s, l))) return rc;
        if ((rc = launch_quadtree(g, h->db, n, s, l))) return rc;
        ORB_CUDA_TRY(cudaEventRecord(h->ev_lvl_done[l], s));
    }
    ORB_CUDA_TRY(cudaEventRecord(h->ev_pyr, st));
    ORB_CUDA_TRY(cudaStreamWaitEvent(h->stream2, h->ev_pyr, 0));
    if ((rc = launch_blur(g, h->db, h->maps_blur, n, h->stream2))) return rc;
    ORB_CUDA_TRY(cudaEventRecord(h->ev_blur, h->stream2));
    for (int l = 0; l < g.nlevels; ++l) ORB_CUDA_TRY(cudaStreamWaitEvent(st, h->ev_lvl_done[l], 0));
    ORB_CUDA_TRY(cudaStreamWaitEvent(st, h->ev_blur, 0));
    return launch_describe(g, h->db, fs, n, st);
}

int launch_sequence(orbx_extractor* h, const FrameSet& fs, int n, cudaStream_t st) {
    const Geometry& g = h->hg;
    int rc;
    if (n <= kLevelParallelMaxBatch) return launch_sequence_level_parallel(h, fs, n, st);
    // the kernels of the main chain are programmatic dependents of each other: the next grid is scheduled while the last
    // blocks of the one before it still run and waits (griddepcontrol.wait) before it reads what that one writes
    for (int l = 1; l < g.nlevels; ++l)
        if ((rc = launch_resize_level(g, h->db, h->maps_resize, l, n, st, l > 1))) return rc;
    // fork: the Gaussian blur (1085-1086) only depends on the pyramid
    ORB_CUDA_TRY(cudaEventRecord(h->ev_pyr, st));
    ORB_CUDA_TRY(cudaStreamWaitEvent(h->stream2, h->ev_pyr, 0));
    if ((rc = launch_blur(g, h->db, h->maps_blur, n, h->stream2))) return rc;
    ORB_CUDA_TRY(cudaEventRecord(h->ev_blur, h->stream2));
    if ((rc = launch_fast(g, h->db, h->maps_fast, n, st, -1, g.nlevels > 1))) return rc;
    if ((rc = launch_quadtree(g, h->db, n, st))) return rc;
    ORB_CUDA_TRY(cudaStreamWaitEvent(st, h->ev_blur, 0));
    return launch_describe(g, h->db, fs, n, st, true);
}

// the whole of operator() for n frames, enqueued on st
int enqueue_pipeline(orbx_extractor* h, const FrameSet& fs, int n, cudaStream_t st) {
    const Geometry& g = h->hg;
    int rc;
    if (fs.base != h->maps_l0.base || fs.pitch != h->maps_l0.pitch || fs.frame_stride != h->maps_l0.frame_stride) {
        const LevelGeom& L = g.lv[0];
        // dims.z = max_batch: the frames of one call are always within the first n <= max_batch slices
        if ((rc = tma_encode_u8_3d(&h->maps_resize.m[0], fs.base, L.w, L.h, h->max_batch, fs.pitch, fs.frame_stride, g.rs_bw, g.rs_bh)) ||
            (rc = tma_encode_u8_3d(&h->maps_blur.m[0], fs.base, L.w, L.h, h->max_batch, fs.pitch, fs.frame_stride, 160, 64)) ||
            (rc = tma_encode_u8_3d(&h->maps_fast.m[0], fs.base, L.w, L.h, h->max_batch, fs.pitch, fs.frame_stride, g.fast_bw, g.max_th)))
            return rc;
        h->maps_l0 = fs;
    }
    if (h->stage_timing) {  // serialised on one stream so that every stage is timed alone
        ORB_CUDA_TRY(cudaEventRecord(h->ev_stage[0], st));
        for (int l = 1; l < g.nlevels; ++l)
            if ((rc = launch_resize_level(g, h->db, h->maps_resize, l, n, st))) return rc;
        ORB_CUDA_TRY(cudaEventRecord(h->ev_stage[1], st));
        if ((rc = launch_blur(g, h->db, h->maps_blur, n, st))) return rc;
        ORB_CUDA_TRY(cudaEventRecord(h->ev_stage[2], st));
        if ((rc = launch_fast(g, h->db, h->maps_fast, n, st))) return rc;
        ORB_CUDA_TRY(cudaEventRecord(h->ev_stage[3], st));
        if ((rc = launch_quadtree(g, h->db, n, st))) return rc;
        ORB_CUDA_TRY(cudaEventRecord(h->ev_stage[4], st));
        if ((rc = launch_describe(g, h->db, fs, n, st))) return rc;
        ORB_CUDA_TRY(cudaEventRecord(h->ev_stage[5], st));
        h->last = fs; h->last_n = n;
        return ORB_OK;
    }
    // small batches: replay the captured graph of the kernel sequence
    constexpr int kGraphMaxBatch = 16;
    const bool capturable = h->use_graphs && n <= kGraphMaxBatch && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
    if (capturable) {
        orbx_extractor::GraphEntry* ge = nullptr;
        orbx_extractor::GraphEntry* lru = &h->graphs[0];
        for (auto& e : h->graphs) {
            if (e.state && e.n == n && e.fs.base == fs.base && e.fs.pitch == fs.pitch && e.fs.frame_stride == fs.frame_stride) ge = &e;
            if (e.stamp < lru->stamp) lru = &e;
        }
        if (ge) ge->stamp = ++h->graph_clock;
        if (ge && ge->state == 2) {
            ORB_CUDA_TRY(cudaGraphLaunch(ge->exec, st));
            count_launch(ge->kernels);
            h->last = fs; h->last_n = n;
            return ORB_OK;
        }
        if (ge && ge->state == 1) {  // second identical call: capture
            const long long before = orb_launch_count();
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                rc = launch_sequence(h, fs, n, st);
                e = cudaStreamEndCapture(st, &graph);
                const int captured = (int)(orb_launch_count() - before);
                count_launch(-captured);  // nothing ran during capture
                if (rc == ORB_OK && e == cudaSuccess && graph) {
                    e = cudaGraphInstantiate(&ge->exec, graph, 0);
                    cudaGraphDestroy(graph);
                    if (e == cudaSuccess) {
                        ge->state = 2; ge->kernels = captured;
                        h->graph_failures = 0;
                        ORB_CUDA_TRY(cudaGraphLaunch(ge->exec, st));
                        count_launch(captured);
                        h->last = fs; h->last_n = n;
                        return ORB_OK;
                    }
                    ge->exec = nullptr;
                } else if (graph) {
                    cudaGraphDestroy(graph);
                }
            }
            cudaGetLastError();      // capture not possible right now (e.g. the caller's stream is already capturing):
            ge->state = 0;           // this call runs as plain launches; only repeated failures switch graphs off
            if (++h->graph_failures >= 3) h->use_graphs = false;
        } else if (!ge) {
            if (lru->exec) { cudaGraphExecDestroy(lru->exec); lru->exec = nullptr; }
            lru->fs = fs; lru->n = n; lru->state = 1; lru->kernels = 0; lru->stamp = ++h->graph_clock;
        }
    }
    if ((rc = launch_sequence(h, fs, n, st))) return rc;
    h->last = fs; h->last_n = n;
    return ORB_OK;
}

}  // namespace

extern "C" {

int orbx_create(const orbx_config* cfg, int device, int width, int height, int max_batch, orbx_handle* out) {
    ORB_REQUIRE(cfg && out, "null pointer");
    ORB_REQUIRE(cfg->nlevels >= 1 && cfg->nlevels <= kMaxLevels, "nlevels out of range (1..16)");
    ORB_REQUIRE(cfg->nfeatures >= 1 && cfg->scale_factor > 1.0f && cfg->scale_factor <= 1.75f, "nfeatures >= 1 and 1 < scale_factor <= 1.75 required");
    ORB_REQUIRE(cfg->min_th_fast >= 1 && cfg->ini_th_fast >= cfg->min_th_fast && cfg->ini_th_fast < 255, "FAST thresholds: 1 <= min <= ini < 255");
    ORB_REQUIRE(width > 0 && height > 0 && max_batch >= 1, "image size / batch must be positive");
    *out = nullptr;
    ORB_CUDA_TRY(cudaSetDevice(device));
    orbx_extractor* h = new (std::nothrow) orbx_extractor();
    ORB_REQUIRE(h, "out of host memory");
    h->cfg = *cfg; h->device = device; h->width = width; h->height = height; h->max_batch = max_batch;
    const int rc = build_geometry(h);
    if (rc != ORB_OK) { orbx_destroy(h); return rc; }
    *out = h;
    return ORB_OK;
}

void orbx_destroy(orbx_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->stream2) cudaStreamSynchronize(h->stream2);
    void* ptrs[] = {h->d_geom, h->d_cells, h->d_taps, h->d_tiles, h->d_pattern, h->db.pyr, h->db.blur, h->db.slots, h->db.cell_counts,
                    h->db.sortbuf, h->db.selected, h->db.sel_counts, h->d_results, h->d_input, h->d_stereo};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->h_pyr) cudaFreeHost(h->h_pyr);
    if (h->h_out) cudaFreeHost(h->h_out);
    for (auto& e : h->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
    for (cudaEvent_t e : h->ev_stage) if (e) cudaEventDestroy(e);
    if (h->ev_pyr) cudaEventDestroy(h->ev_pyr);
    if (h->ev_blur) cudaEventDestroy(h->ev_blur);
    for (int l = 0; l < kMaxLevels; ++l) {
        if (h->lvl_stream[l]) { cudaStreamSynchronize(h->lvl_stream[l]); cudaStreamDestroy(h->lvl_stream[l]); }
        if (h->ev_lvl_ready[l]) cudaEventDestroy(h->ev_lvl_ready[l]);
        if (h->ev_lvl_done[l]) cudaEventDestroy(h->ev_lvl_done[l]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    delete h;
}

int orbx_get_tables(orbx_handle h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* fpl) {
    ORB_REQUIRE(h, "null handle");
    for (int i = 0; i < h->cfg.nlevels; ++i) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->inv_scale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
        if (fpl) fpl[i] = h->quota[i];
    }
    return ORB_OK;
}

int orbx_set_stage_timing(orbx_handle h, int enable) {
    ORB_REQUIRE(h, "null handle");
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    if (enable && !h->ev_stage[0])
        for (int i = 0; i <= ORBX_NUM_STAGES; ++i) ORB_CUDA_TRY(cudaEventCreate(&h->ev_stage[i]));
    h->stage_timing = enable != 0;
    return ORB_OK;
}

int orbx_stage_times(orbx_handle h, float* ms) {
    ORB_REQUIRE(h && ms && h->ev_stage[0], "stage timing was never enabled");
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    ORB_CUDA_TRY(cudaEventSynchronize(h->ev_stage[ORBX_NUM_STAGES]));
    for (int i = 0; i < ORBX_NUM_STAGES; ++i) ORB_CUDA_TRY(cudaEventElapsedTime(&ms[i], h->ev_stage[i], h->ev_stage[i + 1]));
    return ORB_OK;
}

int orbx_algorithmic_bytes(orbx_handle h, double* bytes) {
    ORB_REQUIRE(h && bytes, "null pointer");
    const Geometry& g = h->hg;
    double px = 0, pyr_rw = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        px += (double)g.lv[l].w * g.lv[l].h;
        if (l) pyr_rw += (double)g.lv[l - 1].w * g.lv[l - 1].h + (double)g.lv[l].w * g.lv[l].h;
    }
    bytes[0] = pyr_rw;                                  /* resize: read level l-1, write level l */
    bytes[1] = 2 * px;                                  /* blur: read + write every level */
    bytes[2] = px;                                      /* FAST: read every level once */
    bytes[3] = 0;                                       /* quadtree: latency bound, no streaming */
    bytes[4] = (double)g.out_cap * 56;                  /* describe: result records */
    return ORB_OK;
}

int orbx_max_keypoints(orbx_handle h) { return h ? h->hg.out_cap : ORB_EINVAL; }

int orbx_level_size(orbx_handle h, int level, int* width, int* height) {
    ORB_REQUIRE(h && level >= 0 && level < h->cfg.nlevels, "bad handle / level");
    if (width) *width = h->hg.lv[level].w;
    if (height) *height = h->hg.lv[level].h;
    return ORB_OK;
}

int orbx_extract_device(orbx_handle h, const uint8_t* d_images, size_t stride, size_t frame_stride, int n, void* stream) {
    ORB_REQUIRE(h, "null handle");
    ORB_REQUIRE(n >= 0 && n <= h->max_batch, "batch larger than max_batch");
    if (n == 0) return ORB_OK;
    ORB_REQUIRE(d_images && stride >= (size_t)h->width && frame_stride >= stride * (size_t)(h->height - 1) + h->width, "bad frame layout");
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    if (((uintptr_t)d_images & 15) || (stride & 15) || (frame_stride & 15)) {
        // the TMA tile loads need 16-byte aligned rows: copy such frames into the handle's staging area first
        const size_t dev_frame = h->in_pitch * h->height;
        for (int i = 0; i < n; ++i)
            ORB_CUDA_TRY(cudaMemcpy2DAsync(h->d_input + i * dev_frame, h->in_pitch, d_images + i * frame_stride, stride, h->width, h->height,
                                           cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        FrameSet staged{h->d_input, h->in_pitch, dev_frame};
        return enqueue_pipeline(h, staged, n, (cudaStream_t)stream);
    }
    FrameSet fs{d_images, stride, frame_stride};
    return enqueue_pipeline(h, fs, n, (cudaStream_t)stream);
}

int orbx_device_results(orbx_handle h, const orbx_keypoint** d_kps, const uint8_t** d_desc, const int32_t** d_counts, int* cap_dev) {
    ORB_REQUIRE(h, "null handle");
    if (d_kps) *d_kps = h->db.kps;
    if (d_desc) *d_desc = h->db.desc;
    if (d_counts) *d_counts = h->db.counts;
    if (cap_dev) *cap_dev = h->hg.out_cap;
    return ORB_OK;
}

int orbx_upload_frames(orbx_handle h, const uint8_t* images, size_t stride, size_t frame_stride, int n, void* stream) {
    ORB_REQUIRE(h && images, "null pointer");
    ORB_REQUIRE(n >= 0 && n <= h->max_batch && stride >= (size_t)h->width, "bad batch / stride");
    if (n == 0) return ORB_OK;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t dev_frame = h->in_pitch * h->height;
    if ((frame_stride == stride * (size_t)h->height || n == 1) && stride == h->in_pitch && stride == (size_t)h->width) {
        // tightly packed on both sides: one linear copy (a 2-D descriptor with pitch == width is not guaranteed to be collapsed)
        ORB_CUDA_TRY(cudaMemcpyAsync(h->d_input, images, dev_frame * n, cudaMemcpyHostToDevice, st));
    } else if (frame_stride == stride * (size_t)h->height || n == 1) {
        ORB_CUDA_TRY(cudaMemcpy2DAsync(h->d_input, h->in_pitch, images, stride, h->width, (size_t)h->height * n, cudaMemcpyHostToDevice, st));
    } else {
        for (int i = 0; i < n; ++i)
            ORB_CUDA_TRY(cudaMemcpy2DAsync(h->d_input + i * dev_frame, h->in_pitch, images + i * frame_stride, stride, h->width, h->height,
                                           cudaMemcpyHostToDevice, st));
    }
    return ORB_OK;
}

int orbx_extract_staged(orbx_handle h, int n, void* stream) {
    ORB_REQUIRE(h, "null handle");
    ORB_REQUIRE(n >= 0 && n <= h->max_batch, "batch larger than max_batch");
    if (n == 0) return ORB_OK;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    FrameSet fs{h->d_input, h->in_pitch, h->in_pitch * h->height};
    return enqueue_pipeline(h, fs, n, (cudaStream_t)stream);
}

int orbx_download_results(orbx_handle h, int n, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts, void* stream) {
    ORB_REQUIRE(h && counts, "null pointer");
    ORB_REQUIRE(n >= 0 && n <= h->max_batch && cap >= 0 && (cap == 0 || (kps && desc)), "bad batch / buffers");
    if (n == 0) return ORB_OK;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int oc = h->hg.out_cap, take = std::min(cap, oc);
    ORB_CUDA_TRY(cudaMemcpyAsync(counts, h->db.counts, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (take > 0) {
        ORB_CUDA_TRY(cudaMemcpy2DAsync(kps, (size_t)cap * sizeof(orbx_keypoint), h->db.kps, (size_t)oc * sizeof(orbx_keypoint),
                                       (size_t)take * sizeof(orbx_keypoint), n, cudaMemcpyDeviceToHost, st));
        ORB_CUDA_TRY(cudaMemcpy2DAsync(desc, (size_t)cap * 32, h->db.desc, (size_t)oc * 32, (size_t)take * 32, n, cudaMemcpyDeviceToHost, st));
    }
    return ORB_OK;
}

int orbx_extract_batch(orbx_handle h, const uint8_t* images, size_t stride, size_t frame_stride, int n, orbx_keypoint* kps,
                       uint8_t* desc, int cap, int32_t* counts) {
    ORB_REQUIRE(h, "null handle");
    ORB_REQUIRE(n >= 0 && n <= h->max_batch, "batch larger than max_batch");
    if (n == 0) return ORB_OK;
    ORB_REQUIRE(counts, "null counts");
    if (!images) {  // empty image: the reference returns without touching the outputs (1046-1047)
        for (int i = 0; i < n; ++i) counts[i] = 0;
        return ORB_OK;
    }
    ORB_REQUIRE(stride >= (size_t)h->width && cap >= 0 && (cap == 0 || (kps && desc)), "bad stride / output buffers");
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    int rc;
    if ((rc = orbx_upload_frames(h, images, stride, frame_stride, n, st)) || (rc = orbx_extract_staged(h, n, st))) return rc;
    // Results of a small call come back through pinned staging: a copy into the caller's pageable buffers blocks the host
    // and costs ~15 us each (3 of them: 47 us behind the 75 us of kernels of one frame, measured); three asynchronous copies
    // into pinned memory, one synchronisation and a host memcpy of the used entries cost ~12 us.
    const int oc = h->hg.out_cap, take = std::min(cap, oc);
    const size_t kp_row = (size_t)take * sizeof(orbx_keypoint), de_row = (size_t)take * 32;
    const size_t staged_bytes = align_up((size_t)n * 4, 256) + (size_t)n * (kp_row + de_row);
    if (take == oc && n == h->max_batch && h->results_bytes() <= ((size_t)4 << 20)) {
        // the whole handle is in use: counts | kps | desc are one contiguous slab on the device -> one copy
        const size_t bytes = h->results_bytes();
        if (h->h_out_bytes < bytes) {
            if (h->h_out) { cudaStreamSynchronize(st); cudaFreeHost(h->h_out); h->h_out = nullptr; h->h_out_bytes = 0; }
            ORB_CUDA_TRY(cudaMallocHost(&h->h_out, bytes));
            h->h_out_bytes = bytes;
        }
        ORB_CUDA_TRY(cudaMemcpyAsync(h->h_out, h->d_results, bytes, cudaMemcpyDeviceToHost, st));
        ORB_CUDA_TRY(cudaStreamSynchronize(st));
        const int32_t* s_cnt = reinterpret_cast<const int32_t*>(h->h_out);
        const uint8_t* s_kps = h->h_out + h->results_kps_off();
        const uint8_t* s_desc = s_kps + (size_t)n * kp_row;
        for (int i = 0; i < n; ++i) {
            counts[i] = s_cnt[i];
            const size_t c = (size_t)std::max(0, std::min(s_cnt[i], take));
            memcpy(kps + (size_t)i * cap, s_kps + (size_t)i * kp_row, c * sizeof(orbx_keypoint));
            memcpy(desc + (size_t)i * cap * 32, s_desc + (size_t)i * de_row, c * 32);
        }
    } else if (take > 0 && staged_bytes <= ((size_t)4 << 20)) {
        if (h->h_out_bytes < staged_bytes) {
            if (h->h_out) { cudaStreamSynchronize(st); cudaFreeHost(h->h_out); h->h_out = nullptr; h->h_out_bytes = 0; }
            ORB_CUDA_TRY(cudaMallocHost(&h->h_out, staged_bytes));
            h->h_out_bytes = staged_bytes;
        }
        int32_t* s_cnt = reinterpret_cast<int32_t*>(h->h_out);
        uint8_t* s_kps = h->h_out + align_up((size_t)n * 4, 256);
        uint8_t* s_desc = s_kps + (size_t)n * kp_row;
        ORB_CUDA_TRY(cudaMemcpyAsync(s_cnt, h->db.counts, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA_TRY(cudaMemcpy2DAsync(s_kps, kp_row, h->db.kps, (size_t)oc * sizeof(orbx_keypoint), kp_row, n, cudaMemcpyDeviceToHost, st));
        ORB_CUDA_TRY(cudaMemcpy2DAsync(s_desc, de_row, h->db.desc, (size_t)oc * 32, de_row, n, cudaMemcpyDeviceToHost, st));
        ORB_CUDA_TRY(cudaStreamSynchronize(st));
        for (int i = 0; i < n; ++i) {
            counts[i] = s_cnt[i];
            const size_t c = (size_t)std::max(0, std::min(s_cnt[i], take));
            memcpy(kps + (size_t)i * cap, s_kps + (size_t)i * kp_row, c * sizeof(orbx_keypoint));
            memcpy(desc + (size_t)i * cap * 32, s_desc + (size_t)i * de_row, c * 32);
        }
    } else {
        if ((rc = orbx_download_results(h, n, kps, desc, cap, counts, st))) return rc;
        ORB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    for (int i = 0; i < n; ++i)
        if (counts[i] > cap) { set_error("frame %d has %d keypoints, buffer holds %d", i, counts[i], cap); rc = ORB_ECAPACITY; }
    return rc;
}

int orbx_extract(orbx_handle h, const uint8_t* image, size_t stride, orbx_keypoint* kps, uint8_t* desc, int cap, int* n_out) {
    ORB_REQUIRE(h && n_out, "null pointer");
    int32_t cnt = 0;
    const int rc = orbx_extract_batch(h, image, stride, stride * (size_t)h->height, 1, kps, desc, cap, &cnt);
    *n_out = cnt;
    return rc;
}

int orbx_pyramid_level_device(orbx_handle h, int frame, int level, const uint8_t** d_ptr, size_t* pitch) {
    ORB_REQUIRE(h && level >= 0 && level < h->cfg.nlevels && frame >= 0 && frame < h->last_n, "bad handle / level / frame");
    int p;
    const uint8_t* ptr = level_ptr(h->hg, h->last, h->db.pyr, frame, level, &p);
    if (d_ptr) *d_ptr = ptr;
    if (pitch) *pitch = (size_t)p;
    return ORB_OK;
}

int orbx_pyramid_level(orbx_handle h, int frame, int level, uint8_t* dst, size_t dst_stride) {
    const uint8_t* p; size_t pitch;
    const int rc = orbx_pyramid_level_device(h, frame, level, &p, &pitch);
    if (rc) return rc;
    const LevelGeom& L = h->hg.lv[level];
    ORB_REQUIRE(dst && dst_stride >= (size_t)L.w, "bad destination");
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    ORB_CUDA_TRY(cudaMemcpy2D(dst, dst_stride, p, pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return ORB_OK;
}

// mvImagePyramid in one go: levels [first, first + count) of `frame` -> dst[k] (row stride dst_stride[k] >= width of the
// level). The levels are copied into a pinned staging area by `count` asynchronous 2-D copies on the handle's stream, one
// synchronisation, then host memcpys - instead of one synchronous pageable copy per level.
int orbx_pyramid_levels(orbx_handle h, int frame, int first, int count, uint8_t* const* dst, const size_t* dst_stride) {
    ORB_REQUIRE(h && dst && dst_stride && frame >= 0 && frame < h->last_n, "bad handle / frame");
    ORB_REQUIRE(first >= 0 && count >= 0 && first + count <= h->cfg.nlevels, "bad level range");
    if (count == 0) return ORB_OK;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    // levels >= 1 of a frame are one contiguous slab on the device: ONE linear copy into pinned memory (narrow 2-D copies
    // run far below the link rate), then the rows are unpacked on the host; level 0 lives in the caller's frames
    if (!h->h_pyr) ORB_CUDA_TRY(cudaMallocHost(&h->h_pyr, h->hg.pyr_bytes + (size_t)h->hg.lv[0].w * h->hg.lv[0].h));
    uint8_t* h_l0 = h->h_pyr + h->hg.pyr_bytes;
    bool need_slab = false;
    for (int k = 0; k < count; ++k) {
        const LevelGeom& L = h->hg.lv[first + k];
        ORB_REQUIRE(dst[k] && dst_stride[k] >= (size_t)L.w, "bad destination");
        if (first + k == 0) {
            int pitch;
            const uint8_t* src = level_ptr(h->hg, h->last, h->db.pyr, frame, 0, &pitch);
            ORB_CUDA_TRY(cudaMemcpy2DAsync(h_l0, L.w, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost, h->stream));
        } else need_slab = true;
    }
    if (need_slab)
        ORB_CUDA_TRY(cudaMemcpyAsync(h->h_pyr, h->db.pyr + (size_t)frame * h->hg.pyr_bytes, h->hg.pyr_bytes, cudaMemcpyDeviceToHost, h->stream));
    ORB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < count; ++k) {
        const LevelGeom& L = h->hg.lv[first + k];
        const uint8_t* src = first + k == 0 ? h_l0 : h->h_pyr + L.img_off;
        const size_t sp = first + k == 0 ? (size_t)L.w : (size_t)L.pitch;
        for (int y = 0; y < L.h; ++y) memcpy(dst[k] + (size_t)y * dst_stride[k], src + (size_t)y * sp, L.w);
    }
    return ORB_OK;
}

// ---- stereo (Frame::ComputeStereoMatches, src/Frame.cc:466-640) on the results of two handles ----------
static StereoSide stereo_side(orbx_extractor* h, int frame) {
    StereoSide s;
    s.g = h->db.geom; s.fs = h->last; s.pyr = h->db.pyr; s.frame = frame;
    s.kps = h->db.kps + (size_t)frame * h->hg.out_cap;
    s.desc = h->db.desc + (size_t)frame * h->hg.out_cap * 32;
    s.count = h->db.counts + frame;
    return s;
}

int orbm_stereo_match_device(orbx_handle left, orbx_handle right, int frame, float mbf, float mb, float* d_uRight, float* d_depth,
                             int32_t* d_sad, int32_t* d_kept, void* stream) {
    ORB_REQUIRE(left && right && d_uRight && d_depth && d_sad && d_kept, "null pointer");
    ORB_REQUIRE(left->device == right->device && left->width == right->width && left->height == right->height &&
                    left->cfg.nlevels == right->cfg.nlevels && left->cfg.scale_factor == right->cfg.scale_factor,
                "left / right extractors must share device, image size and pyramid settings");
    ORB_REQUIRE(frame >= 0 && frame < left->last_n && frame < right->last_n, "frame was not extracted by both handles");
    ORB_REQUIRE(mb > 0 && mbf > 0, "baseline must be positive");
    ORB_CUDA_TRY(cudaSetDevice(left->device));
    return launch_stereo(stereo_side(left, frame), stereo_side(right, frame), left->hg.out_cap, 1, left->hg.out_cap, mbf, mb, d_uRight, d_depth,
                         d_sad, d_kept, (cudaStream_t)stream);
}

int orbm_stereo_match_batch_device(orbx_handle left, orbx_handle right, int n_frames, float mbf, float mb, float* d_uRight, float* d_depth,
                                   int32_t* d_sad, int32_t* d_kept, int out_stride, void* stream) {
    ORB_REQUIRE(left && right && d_uRight && d_depth && d_sad && d_kept, "null pointer");
    ORB_REQUIRE(left->device == right->device && left->width == right->width && left->height == right->height &&
                    left->cfg.nlevels == right->cfg.nlevels && left->cfg.scale_factor == right->cfg.scale_factor,
                "left / right extractors must share device, image size and pyramid settings");
    ORB_REQUIRE(n_frames >= 1 && n_frames <= left->last_n && n_frames <= right->last_n, "more frames than both handles extracted");
    ORB_REQUIRE(out_stride >= left->hg.out_cap, "out_stride must be at least orbx_max_keypoints(left)");
    ORB_REQUIRE(mb > 0 && mbf > 0, "baseline must be positive");
    ORB_CUDA_TRY(cudaSetDevice(left->device));
    return launch_stereo(stereo_side(left, 0), stereo_side(right, 0), left->hg.out_cap, n_frames, out_stride, mbf, mb, d_uRight, d_depth, d_sad,
                         d_kept, (cudaStream_t)stream);
}

int orbm_stereo_match(orbx_handle left, orbx_handle right, int frame, float mbf, float mb, float* uRight, float* depth, int cap, int* kept) {
    ORB_REQUIRE(left && uRight && depth && kept, "null pointer");
    const int oc = left->hg.out_cap;
    ORB_CUDA_TRY(cudaSetDevice(left->device));
    if (!left->d_stereo) ORB_CUDA_TRY(cudaMalloc(&left->d_stereo, (size_t)oc * 12 + 16));
    float* d = left->d_stereo;
    float* d_u = d; float* d_d = d + oc; int* d_s = (int*)(d + 2 * (size_t)oc); int* d_k = d_s + oc;
    cudaStream_t st = left->stream;
    int rc = orbm_stereo_match_device(left, right, frame, mbf, mb, d_u, d_d, d_s, d_k, st);
    if (rc == ORB_OK) {
        // uRight | depth are adjacent on the device: one copy into pinned staging + the counter, one synchronisation, then
        // host memcpys (copies straight into the caller's pageable arrays block the host one by one, see orbx_extract_batch)
        const int take = std::max(0, std::min(cap, oc));
        const size_t need = (size_t)oc * 8 + 16;
        if (left->h_out_bytes < need) {
            if (left->h_out) { cudaStreamSynchronize(st); cudaFreeHost(left->h_out); left->h_out = nullptr; left->h_out_bytes = 0; }
            ORB_CUDA_TRY(cudaMallocHost(&left->h_out, need));
            left->h_out_bytes = need;
        }
        float* s_u = reinterpret_cast<float*>(left->h_out);
        int* s_k = reinterpret_cast<int*>(left->h_out + (size_t)oc * 8);
        cudaError_t e = cudaMemcpyAsync(s_u, d_u, (size_t)oc * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s_k, d_k, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("stereo copy failed: %s", cudaGetErrorString(e)); rc = ORB_ECUDA; }
        else {
            memcpy(uRight, s_u, (size_t)take * 4);
            memcpy(depth, s_u + oc, (size_t)take * 4);
            *kept = *s_k;
        }
    }
    return rc;
}

int orbx_debug_blurred_level(orbx_handle h, int frame, int level, uint8_t* dst, size_t dst_stride) {
    ORB_REQUIRE(h && level >= 0 && level < h->cfg.nlevels && frame >= 0 && frame < h->last_n, "bad handle / level / frame");
    const LevelGeom& L = h->hg.lv[level];
    ORB_REQUIRE(dst && dst_stride >= (size_t)L.w, "bad destination");
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    ORB_CUDA_TRY(cudaMemcpy2D(dst, dst_stride, h->db.blur + (size_t)frame * h->hg.blur_bytes + L.blur_off, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return ORB_OK;
}

int orbx_debug_candidates(orbx_handle h, int frame, int level, int32_t* xys, int cap, int* n_out) {
    ORB_REQUIRE(h && n_out && level >= 0 && level < h->cfg.nlevels && frame >= 0 && frame < h->last_n, "bad handle / level / frame");
    const Geometry& g = h->hg;
    const LevelGeom& L = g.lv[level];
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    std::vector<int> cnt(L.cell_count);
    std::vector<uint32_t> sl((size_t)L.cand_cap);
    ORB_CUDA_TRY(cudaMemcpy(cnt.data(), h->db.cell_counts + (size_t)frame * g.ncells + L.cell_begin, cnt.size() * 4, cudaMemcpyDeviceToHost));
    ORB_CUDA_TRY(cudaMemcpy(sl.data(), h->db.slots + (size_t)frame * g.slot_words + L.slot_off, sl.size() * 4, cudaMemcpyDeviceToHost));
    int n = 0;
    for (int c = 0; c < L.cell_count; ++c)
        for (int k = 0; k < cnt[c]; ++k, ++n) {
            if (n >= cap) continue;
            const uint32_t p = sl[(size_t)c * L.slot_cap + k];
            xys[3 * n] = (int)(p & 0xfff) + kMinBorder; xys[3 * n + 1] = (int)((p >> 12) & 0xfff) + kMinBorder; xys[3 * n + 2] = (int)(p >> 24);
        }
    *n_out = n;
    return n > cap ? ORB_ECAPACITY : ORB_OK;
}

int orbx_debug_quadtree(int device, const int32_t* xs, const int32_t* ys, const int32_t* scores, int n, int minX, int maxX,
                        int minY, int maxY, int N, int32_t* out_idx, int cap, int* n_out) {
    ORB_REQUIRE(n >= 0 && n_out && (n == 0 || (xs && ys && scores)), "bad arguments");
    *n_out = 0;
    if (n == 0) return ORB_OK;
    const int W = maxX - minX, H = maxY - minY;
    ORB_REQUIRE(W > 0 && H > 0 && W <= 4096 && H <= 4096, "bad bounds");
    const int nRoots = (int)roundf((float)W / (float)H);
    ORB_REQUIRE(nRoots >= 1, "aspect ratio below 0.5");
    const float rootW = (float)W / nRoots;
    const int span = std::max((int)ceilf(rootW) + 1, H);
    int depth = 1;
    while ((1 << depth) < span) ++depth;
    depth = std::min(depth + 1, kQtMaxDepth);
    const int sel_cap = std::max(N, 4 * nRoots) + 3;
    std::vector<uint32_t> packed(n);
    for (int i = 0; i < n; ++i) {
        ORB_REQUIRE(xs[i] >= 0 && xs[i] < 4096 && ys[i] >= 0 && ys[i] < 4096 && scores[i] >= 0 && scores[i] < 256, "candidate out of range");
        packed[i] = (uint32_t)xs[i] | (uint32_t)ys[i] << 12 | (uint32_t)scores[i] << 24;
    }
    ORB_CUDA_TRY(cudaSetDevice(device));
    uint32_t *d_cand = nullptr, *d_scr = nullptr, *d_sel = nullptr; int* d_cnt = nullptr;
    ORB_CUDA_TRY(cudaMalloc(&d_cand, (size_t)n * 4)); ORB_CUDA_TRY(cudaMalloc(&d_scr, (size_t)n * 16));
    ORB_CUDA_TRY(cudaMalloc(&d_sel, (size_t)sel_cap * 4)); ORB_CUDA_TRY(cudaMalloc(&d_cnt, 4));
    ORB_CUDA_TRY(cudaMemcpy(d_cand, packed.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    int rc = launch_quadtree_standalone(d_cand, n, N, nRoots, rootW, H, depth, d_scr, d_sel, sel_cap, d_cnt, 0);
    int cnt = 0;
    std::vector<uint32_t> sel(sel_cap);
    if (!rc) {
        cudaError_t e = cudaMemcpy(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(sel.data(), d_sel, (size_t)sel_cap * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("quadtree kernel failed: %s", cudaGetErrorString(e)); rc = ORB_ECUDA; }
    }
    cudaFree(d_cand); cudaFree(d_scr); cudaFree(d_sel); cudaFree(d_cnt);
    if (rc) return rc;
    // map the packed survivors back to candidate indices (pixels are unique per level)
    for (int k = 0; k < cnt && k < cap; ++k) {
        int found = -1;
        for (int i = 0; i < n; ++i) if (packed[i] == sel[k]) { found = i; break; }
        out_idx[k] = found;
    }
    *n_out = cnt;
    return cnt > cap ? ORB_ECAPACITY : ORB_OK;
}

}  // extern "C"
