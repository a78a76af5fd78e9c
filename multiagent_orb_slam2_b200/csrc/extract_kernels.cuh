// Device-side data layout of the extractor and the kernel launch interface.
//
// Per extractor handle everything lives in a handful of flat HBM allocations, frame-major:
//   pyramid slab   [B][pyr_bytes]      levels 1..L-1 at LevelGeom::img_off, pitch 128-aligned
//                                      (level 0 is read in place from the caller's frames)
//   blurred slab   [B][blur_bytes]     levels 0..L-1
//   cell slots     [B][slot_words]     per FAST cell a fixed-capacity list of packed candidates
//   cell counts    [B][ncells]
//   sort scratch   [B][5*cand_words]   dense candidates + ping-pong (key, value) arrays
//   selected       [B][sel_words]      per level the quadtree survivors, list order
//   sel counts     [B][L]
//   results        [B][cap] keypoints, [B][cap][32] descriptors, [B] counts
// A packed candidate is x | y << 12 | score << 24 with x, y relative to the (16,16) border like the
// reference's vToDistributeKeys (src/ORBextractor.cc:820-825).
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace orb {

constexpr int kMaxLevels = 16;
constexpr int kEdgeThreshold = 19;
constexpr int kMinBorder = kEdgeThreshold - 3;  // 16
constexpr int kHalfPatch = 15;
constexpr int kQtMaxDepth = 13;
// IC_Angle by word loads: 31 rows x 9 words (31 + 3 alignment bytes) = 279 items, padded to 9 warp steps; 4 phases x (u, v)
constexpr int kIcWordsPerRow = 9, kIcItems = 288, kIcTableWords = 4 * kIcItems;   // [alignment phase][item]: u + 16 of the word's 4 bytes, 0 outside the patch
constexpr int kPatternWords = 256;   // [test of the byte][descriptor byte]: x0 | y0 << 8 | x1 << 16 | y1 << 24 as int8

struct LevelGeom {
    int w, h, pitch;
    int nCols, nRows, wCell, hCell;
    int cell_begin, cell_count;  // range in the cell table
    int slot_cap;                // candidates per cell
    int cand_cap;                // cell_count * slot_cap
    int quota;                   // N of DistributeOctTree
    int nRoots;
    float rootW;
    int key_depth;               // path digits kept in the sort key
    int sel_cap, sel_off;        // capacity / offset in the per-frame selected array
    int tab_x_off, tab_y_off;    // resize tables (level >= 1), index into the LinTap array
    int patch_size;              // int(31 * scale)
    float scale;
    size_t img_off;              // in the pyramid slab (levels >= 1)
    size_t blur_off;             // in the blurred slab
    size_t slot_off;             // words, in the per-frame slot array
    size_t cand_off;             // words, in one of the 5 per-frame sort arrays
};

struct CellDesc {
    int16_t level, x0, y0, tw, th;  // tile origin (level coords) and size incl. the 6 px halo
    int16_t offx, offy;             // j*wCell, i*hCell added to cell-local keypoints
    int16_t pad;
    int32_t ordinal;                // position among the level's active cells (row-major)
    uint32_t ginv;                  // ceil(65536 / G), G = aligned 4-byte groups covering the interior columns (fast.cu)
};

struct LinTap { int ofs; short c0, c1; };  // cv::resize INTER_LINEAR tap: source index + 2048-scaled weights

struct BlurTile { int16_t level, tx, ty, pad; };

struct Geometry {  // device copy, read by every kernel
    int nlevels, ncells, ntiles;
    int iniTh, minTh;
    int max_tw, max_th;  // largest FAST tile
    int fast_bw;         // TMA box of the FAST tiles: fast_bw x max_th bytes (fast_bw = row pitch in smem)
    int rs_bw, rs_bh;    // TMA box of the resize source window
    int sel_words;       // per-frame selected capacity (sum of sel_cap)
    int out_cap;         // per-frame result capacity
    size_t pyr_bytes, blur_bytes, slot_words, cand_words;
    LevelGeom lv[kMaxLevels];
    int umax[16];
};

struct FrameSet {  // where the frames of this call live
    const uint8_t* base;   // level 0
    size_t pitch, frame_stride;
};

struct DeviceBuffers {
    const Geometry* geom;  // device pointer
    const CellDesc* cells;
    const LinTap* taps;
    const BlurTile* tiles;
    const uint32_t* ic_table;  // kIcTableWords
    const uint32_t* pattern;   // kPatternWords
    uint8_t* pyr;
    uint8_t* blur;
    uint32_t* slots;
    int* cell_counts;
    uint32_t* sortbuf;
    uint32_t* selected;
    int* sel_counts;
    orbx_keypoint* kps;
    uint8_t* desc;
    int* counts;
};

struct StereoSide {  // one image of a stereo pair: geometry, pyramid and extraction results of a frame
    const Geometry* g;
    FrameSet fs;
    const uint8_t* pyr;
    const orbx_keypoint* kps;
    const uint8_t* desc;
    const int* count;
    int frame;
};
int launch_stereo(const StereoSide& L, const StereoSide& R, int capL, int n_frames, int out_stride, float mbf, float mb, float* d_uRight,
                  float* d_depth, int* d_sad, int* d_kept, cudaStream_t st);

// launchers (each enqueues on `st`; n = frames in this call)
int launch_resize_level(const Geometry& hg, const DeviceBuffers& db, const TmaMaps& maps, int level, int n, cudaStream_t st, bool pdl = false);
int launch_blur(const Geometry& hg, const DeviceBuffers& db, const TmaMaps& maps, int n, cudaStream_t st);
// pdl: launched as a programmatic dependent of the previous kernel in the stream (common.cuh)
int launch_fast(const Geometry& hg, const DeviceBuffers& db, const TmaMaps& maps, int n, cudaStream_t st, int level = -1, bool pdl = false);
int launch_quadtree(const Geometry& hg, const DeviceBuffers& db, int n, cudaStream_t st, int level = -1);
int launch_describe(const Geometry& hg, const DeviceBuffers& db, const FrameSet& fs, int n, cudaStream_t st, bool pdl = false);
int launch_quadtree_standalone(const uint32_t* d_cand, int n, int N, int nRoots, float rootW, int H, int key_depth,
                               uint32_t* d_scratch4n, uint32_t* d_sel, int sel_cap, int* d_count, cudaStream_t st);

__host__ __device__ inline const uint8_t* level_ptr(const Geometry& g, const FrameSet& fs, const uint8_t* pyr, int frame,
                                                    int level, int* pitch) {
    if (level == 0) { *pitch = (int)fs.pitch; return fs.base + (size_t)frame * fs.frame_stride; }
    *pitch = g.lv[level].pitch;
    return pyr + (size_t)frame * g.pyr_bytes + g.lv[level].img_off;
}

}  // namespace orb
