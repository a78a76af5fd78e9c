/* orb_b200.h - C ABI of the B200-native ORB frontend (extraction + binary descriptor matching).
 *
 * Drop-in boundary for the data-parallel hot path of andresenwc/MultiAgent_ORB_SLAM2:
 *   ORB_SLAM2::ORBextractor   include/ORBextractor.h:45-111, src/ORBextractor.cc:410-1132
 *   ORB_SLAM2::ORBmatcher     include/ORBmatcher.h:37-102,   src/ORBmatcher.cc:37-1665
 *   Frame::ComputeStereoMatches                               src/Frame.cc:466-640
 * The C++ facade classes with the reference's own signatures live in include/orbslam2_b200/ and call
 * only the functions below. Plain pointers and sizes; no C++/torch types; never throws.
 *
 * All entry points return ORB_OK (0) or a negative ORB_E* code. There is no CPU fallback: without a
 * CUDA device every compute call returns ORB_ECUDA.
 *
 * Streams: `stream` arguments are cudaStream_t values passed as void* (NULL = the legacy default
 * stream). "_device" functions take device pointers and only enqueue work on `stream`;
 * functions without the suffix take HOST pointers, copy in/out and synchronise before returning.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORB_OK = 0,
    ORB_EINVAL = -1,   /* bad argument (null pointer, size out of range, image too small/large) */
    ORB_ECUDA = -2,    /* CUDA runtime error; orb_last_error() has the text */
    ORB_ECAPACITY = -3 /* result does not fit the caller's buffer / handle capacity */
};

const char* orb_last_error(void);          /* thread-local text of the last failure */
int orb_device_count(void);                /* number of visible CUDA devices (0 = none) */
const char* orb_version(void);
long long orb_launch_count(void);          /* kernels this library has launched so far (process-wide) */

/* ------------------------------------------------------------------------------------------
 * Extractor. Replaces ORBextractor::ORBextractor (src/ORBextractor.cc:410-470) and
 * ORBextractor::operator() (1043-1105) incl. ComputePyramid (1107-1132),
 * ComputeKeyPointsOctTree (765-853), DistributeOctTree (539-763), IC_Angle (77-104) and
 * computeOrbDescriptor (108-147).
 * One handle == one reference ORBextractor instance: not re-entrant, distinct handles may run
 * concurrently (src/Frame.cc:78-81). A handle is pinned to one device (one agent per GPU). */
typedef struct orbx_extractor* orbx_handle;

typedef struct {
    int nfeatures;      /* ORBextractor.nFeatures   */
    float scale_factor; /* ORBextractor.scaleFactor */
    int nlevels;        /* ORBextractor.nLevels (1..16) */
    int ini_th_fast;    /* ORBextractor.iniThFAST   */
    int min_th_fast;    /* ORBextractor.minThFAST   */
} orbx_config;

/* One keypoint = cv::KeyPoint minus class_id (always -1 in the reference). 24 bytes. */
typedef struct {
    float x, y;     /* pt, already multiplied by the level scale (src/ORBextractor.cc:1095-1101) */
    float size;     /* int(31 * scale[octave])  (837-846) */
    float angle;    /* degrees [0,360), cv::fastAtan2 of the intensity centroid */
    float response; /* FAST score */
    int32_t octave; /* pyramid level */
} orbx_keypoint;

/* Image geometry is fixed per handle (all frames of an agent share a camera). max_batch frames can
 * be extracted per call. */
int orbx_create(const orbx_config* cfg, int device, int width, int height, int max_batch, orbx_handle* out);
void orbx_destroy(orbx_handle h);

/* Tables of the ctor, for the getters of the facade (GetScaleFactors() etc., ORBextractor.h:63-83).
 * Each array has nlevels entries; any pointer may be NULL. */
int orbx_get_tables(orbx_handle h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int32_t* features_per_level);
/* Upper bound of keypoints per frame (nfeatures + per-level overshoot, src/ORBextractor.cc:669-737). */
int orbx_max_keypoints(orbx_handle h);
int orbx_level_size(orbx_handle h, int level, int* width, int* height);

/* operator(): one frame from host memory. Writes min(n, cap) keypoints + 32-byte descriptors in the
 * reference's output order (levels 0..n-1, final node-list order inside a level) and the true count
 * to *n_out. Returns ORB_ECAPACITY (after filling cap entries) when n > cap. An empty image
 * (NULL / zero size) is the reference's silent no-op: ORB_OK with *n_out = 0. */
int orbx_extract(orbx_handle h, const uint8_t* image, size_t stride, orbx_keypoint* kps, uint8_t* desc,
                 int cap, int* n_out);

/* Batched operator() over `n` host frames (n <= max_batch), frame i at images + i*frame_stride.
 * Outputs are [n][cap] arrays; counts[n]. */
int orbx_extract_batch(orbx_handle h, const uint8_t* images, size_t stride, size_t frame_stride, int n,
                       orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts);

/* Device-resident variant: frames already in device memory; only enqueues on `stream`. Results stay
 * on the device: d_kps[n][cap_dev], d_desc[n][cap_dev][32], d_counts[n] with cap_dev = orbx_max_keypoints(). */
int orbx_extract_device(orbx_handle h, const uint8_t* d_images, size_t stride, size_t frame_stride, int n,
                        void* stream);
int orbx_device_results(orbx_handle h, const orbx_keypoint** d_kps, const uint8_t** d_desc,
                        const int32_t** d_counts, int* cap_dev);

/* Asynchronous building blocks of orbx_extract_batch for callers that pipeline copies and compute
 * themselves (pinned host memory recommended). All three only enqueue on `stream`:
 *   upload   : host frames -> the handle's device staging area
 *   staged   : operator() over the n staged frames
 *   download : counts[n], kps[n][cap], desc[n][cap][32] -> host (cap <= orbx_max_keypoints()) */
int orbx_upload_frames(orbx_handle h, const uint8_t* images, size_t stride, size_t frame_stride, int n, void* stream);
int orbx_extract_staged(orbx_handle h, int n, void* stream);
int orbx_download_results(orbx_handle h, int n, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts, void* stream);

/* mvImagePyramid (ORBextractor.h:85): level `level` of frame `frame` of the last call. Device view
 * (pointer + pitch) or a copy to host (dst_stride >= width). */
int orbx_pyramid_level_device(orbx_handle h, int frame, int level, const uint8_t** d_ptr, size_t* pitch);
int orbx_pyramid_level(orbx_handle h, int frame, int level, uint8_t* dst, size_t dst_stride);
/* Levels [first, first + count) at once: dst[k] / dst_stride[k] for level first + k. One synchronisation for all of them
 * (pinned staging inside the handle) instead of one synchronous pageable copy per level; level 0 is the caller's own image
 * (ComputePyramid copies it, src/ORBextractor.cc:1107-1132), so a host caller only needs levels 1 .. nlevels-1. */
int orbx_pyramid_levels(orbx_handle h, int frame, int first, int count, uint8_t* const* dst, const size_t* dst_stride);

/* ------------------------------------------------------------------------------------------
 * Matcher primitives. Replace the DescriptorDistance loops of ORBmatcher (src/ORBmatcher.cc:1649-1665
 * called from 102, 216, 387, 444, 586, 738, 944, 1074, 1214, 1294, 1419, 1548; src/Frame.cc:541).
 * Descriptors are 32 bytes, rows contiguous. Selection rule of every reference search: strict '<'
 * updates in iteration order => (two smallest distances, index of the FIRST minimum); both
 * distances start at 256; idx = -1 when no candidate is closer than 256.
 * The brute-force searches have two implementations with identical results - the POPC kernel (csrc/hamming.cu) and the
 * tensor-core kernel (csrc/hamming_mma.cu: +-1 int8 expansion, tcgen05.mma kind::i8, top-2 epilogue out of TMEM) -
 * chosen by problem size; orb_b200_debug.h can pin one of them for the calling thread (tests, measurements). */

/* Brute force: every row of A against every row of B (SearchByBoW inner loop with the gate removed,
 * src/ORBmatcher.cc:566-598; BASELINE configs 3 and 5). */
int orbm_knn2_device(const uint8_t* dA, int nA, const uint8_t* dB, int nB, int32_t* d_idx, int32_t* d_best,
                     int32_t* d_second, void* stream);
int orbm_knn2(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, int32_t* idx, int32_t* best,
              int32_t* second);

/* The tensor-core implementation keeps its expanded operands (256 bytes per descriptor row of the largest call so far) in
 * scratch memory per (device, stream). A device-call user releases it for a stream it is about to destroy, or to return
 * the memory; stream-ordered, no synchronisation. The host-buffer entry points do this for their own stream. */
int orbm_release_scratch(void* stream);

/* Batched brute force: pair p matches A[p] (nA[p] rows at dA + p*strideA_rows*32) against B[p].
 * d_nA / d_nB are device int arrays (e.g. the d_counts of two extractor runs). Outputs [pairs][strideA_rows]. */
int orbm_knn2_batched_device(const uint8_t* dA, const int32_t* d_nA, int strideA_rows, const uint8_t* dB,
                             const int32_t* d_nB, int strideB_rows, int pairs, int32_t* d_idx, int32_t* d_best,
                             int32_t* d_second, void* stream);

/* Same over a list of (query set, database set) index pairs into ONE array of descriptor sets
 * (set k at d_sets + k*stride_rows*32, d_counts[k] rows): frame-to-frame matching of a batch of frames,
 * or every map against every other map (MapFusion). d_pairs = int32 [pairs][2] on the device. */
int orbm_knn2_pairs_device(const uint8_t* d_sets, const int32_t* d_counts, int stride_rows, const int32_t* d_pairs,
                           int pairs, int32_t* d_idx, int32_t* d_best, int32_t* d_second, void* stream);

/* Candidate-list search: query i is compared with B[cands[offsets[i] .. offsets[i+1])] in list order
 * (the GetFeaturesInArea / BoW-node gated loops, e.g. src/ORBmatcher.cc:427-459, 1385-1426). */
int orbm_knn2_lists_device(const uint8_t* dA, int nA, const uint8_t* dB, const int32_t* d_offsets,
                           const int32_t* d_cands, int32_t* d_idx, int32_t* d_best, int32_t* d_second, void* stream);
int orbm_knn2_lists(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, const int32_t* offsets,
                    const int32_t* cands, int32_t* idx, int32_t* best, int32_t* second);

/* Raw distances of the same CSR lists, out[k] = distance(A[q], B[cands[k]]) for k in [offsets[q], offsets[q+1]).
 * For the STATEFUL searches (SearchForInitialization's vMatchedDistance gate src/ORBmatcher.cc:446, the
 * "already matched" flags at 211 / 578): the device does the popcounts, the caller replays the reference's
 * loop over the precomputed distances in the reference's order, which keeps match sets bit-exact. */
int orbm_list_distances_device(const uint8_t* dA, int nA, const uint8_t* dB, const int32_t* d_offsets,
                               const int32_t* d_cands, int16_t* d_out, void* stream);
int orbm_list_distances(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, const int32_t* offsets,
                        const int32_t* cands, int16_t* out);

/* Acceptance test of SearchByBoW(KF,KF) (src/ORBmatcher.cc:600-603): match[i] = idx[i] if
 * best < th_strict_upper && (float)best < ratio*(float)second, else -1. With inclusive != 0 the
 * threshold test is best <= th (SearchForInitialization / SearchByBoW(KF,F), 229-232, 461-463). */
int orbm_ratio_filter_device(const int32_t* d_idx, const int32_t* d_best, const int32_t* d_second, int n,
                             int th, int inclusive, float ratio, int32_t* d_match, void* stream);

/* Stereo matching of a rectified pair: replaces Frame::ComputeStereoMatches (src/Frame.cc:466-640) for
 * frame `frame` of the last calls of two extractor handles (left, right; same device and geometry).
 * Row-band candidate gate, best Hamming < TH_HIGH, 11x11 L1 patch refinement on the left keypoint's
 * pyramid level (read in place from the handles' device pyramids), parabola sub-pixel fit, and the
 * final 1.5*1.4*median rejection. mbf = baseline*fx, mb = baseline (mbf/mb = max disparity).
 * Outputs for the left keypoints i < count: uRight[i], depth[i] (-1 = no match), the patch distance
 * sad[i] (-1 = none) and *kept = number of surviving matches. The row index of mvuRight/mvDepth is the
 * left keypoint index, as in the reference. */
int orbm_stereo_match_device(orbx_handle left, orbx_handle right, int frame, float mbf, float mb, float* d_uRight,
                             float* d_depth, int32_t* d_sad, int32_t* d_kept, void* stream);
int orbm_stereo_match(orbx_handle left, orbx_handle right, int frame, float mbf, float mb, float* uRight,
                      float* depth, int cap, int* kept);
/* The same for frames 0 .. n_frames-1 of the last (batched) calls of both handles in ONE launch pair (BASELINE config 2
 * as a batch): outputs [n_frames][out_stride] (out_stride >= orbx_max_keypoints(left)), d_kept[n_frames]. */
int orbm_stereo_match_batch_device(orbx_handle left, orbx_handle right, int n_frames, float mbf, float mb, float* d_uRight,
                                   float* d_depth, int32_t* d_sad, int32_t* d_kept, int out_stride, void* stream);

/* Full distance matrix (nA x nB, int16) for the ordered greedy resolve of the stateful searches and for
 * MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:246-311). */
int orbm_distance_matrix_device(const uint8_t* dA, int nA, const uint8_t* dB, int nB, int16_t* d_out, void* stream);
int orbm_distance_matrix(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, int16_t* out);

/* ------------------------------------------------------------------------------------------
 * Cross-map matching over the GPUs of one node (BASELINE config 5; SURVEY.md section 8b "orbm_hamming_knn2_allgather", 8e).
 * Replaces, for whole maps, what MapFusion::ComputeSim3 / CovisibilityDiscovery do keyframe by keyframe through
 * ORBmatcher::SearchByBoW(KF, KF) (src/MapFusion.cc:275, 849 -> src/ORBmatcher.cc:524-657): every descriptor of map a
 * against every descriptor of map b != a, (first-minimum index, best, second) per query row.
 * n_maps maps live on `world` ranks (GPUs): map m on rank m % world (its slot-th map there, slot = m / world). The
 * exchange is part of the call: no collective library runs on the data path. Every rank owns a window of device memory
 * that its peers map (CUDA IPC between processes, peer access inside one process); a rank reads the packed descriptor
 * sets it needs straight out of the owners' windows over NVLink inside the kernel that expands them into the tensor-core
 * operands, and writes results into the window of the query map's owner. The directed pairs are cut into 128-row query
 * tiles and dealt evenly to the ranks, so fewer maps than GPUs still keep every GPU busy (query-row split); a row's
 * result is always computed over the whole database in canonical order, i.e. it is identical to the single-GPU result.
 * Protocol per step (one orbm_knn2_allgather call on EVERY rank, same n_maps / rows_per_map everywhere): publish owned sets
 * (release flag) -> acquire peers' flags as their sets are first needed -> match -> release "done" into every window ->
 * owners acquire all contributors. Waits are bounded (seconds) and trap rather than hang.
 *   orbm_xmap_create       window + operand scratch for maps of up to rows_cap descriptors
 *   orbm_xmap_ipc_handle   64 opaque bytes (cudaIpcMemHandle_t) to hand to the other processes by any host transport
 *   orbm_xmap_attach_ipc   handles of all ranks, rank order (own entry ignored); multi-process
 *   orbm_xmap_attach_local contexts of all ranks created in THIS process (one thread driving several GPUs, as the
 *                          reference's single-process MultiAgentServer would); enables peer access
 *   orbm_knn2_allgather    d_sets[k] = device pointer to the k-th map this rank owns (rows_per_map[rank + k*world] x 32 B);
 *                          rows_per_map = host array of all n_maps counts; only enqueues on `stream`
 *   orbm_xmap_result       device pointers (valid after the stream has finished the step) to idx / best / second of pair
 *                          (map_a, map_b), rows_per_map[map_a] entries each, on the rank that owns map_a
 *   orbm_xmap_plan         host only: the (a, b, first tile, end tile) chunks of `rank` - the split itself, for tests */
#define ORBM_XMAP_HANDLE_BYTES 64
typedef struct orbm_xmap* orbm_xmap_t;
int orbm_xmap_create(int device, int rank, int world, int n_maps, int rows_cap, orbm_xmap_t* out);
void orbm_xmap_destroy(orbm_xmap_t x);
int orbm_xmap_ipc_handle(orbm_xmap_t x, void* handle64);
int orbm_xmap_attach_ipc(orbm_xmap_t x, const void* handles /* world x ORBM_XMAP_HANDLE_BYTES */);
int orbm_xmap_attach_local(orbm_xmap_t* ctxs, int world);
int orbm_knn2_allgather(orbm_xmap_t x, const uint8_t* const* d_sets, const int32_t* rows_per_map, void* stream);
int orbm_xmap_result(orbm_xmap_t x, int map_a, int map_b, const int32_t** d_idx, const int32_t** d_best, const int32_t** d_second);
int orbm_xmap_plan(int n_maps, const int32_t* rows_per_map, int world, int rank, int32_t* chunks /* [cap][4] */, int cap, int* n_chunks);

/* ------------------------------------------------------------------------------------------
 * Device-resident Frame grid + windowed search. orbm_grid_build_device replaces
 * Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:230-245, 382-392) on the keypoints an extractor left on the
 * device (bounds = mnMinX, mnMinY, mnMaxX, mnMaxY); orbm_window_knn2_device replaces Frame::GetFeaturesInArea
 * (src/Frame.cc:327-380) fused with the best / second loop that consumes it (e.g. src/ORBmatcher.cc:84-116): query q
 * = (x, y, r, minLevel, maxLevel, descriptor row q of d_queries); the grid and the window tests use the coordinates of the
 * keypoint array given to orbm_grid_build_device - for a distorted camera pass mvKeysUn (orbm_undistort_keypoints_device)
 * and the bounds of orbm_image_bounds; candidates are visited in the reference's order
 * (cell column, cell row, insertion), so ties resolve identically. Outputs also carry the octaves of best and second
 * (src/ORBmatcher.cc:119-122). Stateless form; capacity < 2^20 keypoints per frame. */
typedef struct orbm_grid* orbm_grid_handle;
int orbm_grid_create(int device, int capacity, orbm_grid_handle* out);
void orbm_grid_destroy(orbm_grid_handle g);
int orbm_grid_build_device(orbm_grid_handle g, const orbx_keypoint* d_kps, const int32_t* d_count, float min_x, float min_y,
                           float max_x, float max_y, void* stream);
int orbm_window_knn2_device(orbm_grid_handle g, const uint8_t* d_desc_frame, const uint8_t* d_queries, int nq, const float* d_x,
                            const float* d_y, const float* d_r, const int32_t* d_min_level, const int32_t* d_max_level,
                            int32_t* d_idx, int32_t* d_best, int32_t* d_second, int32_t* d_best_level, int32_t* d_second_level,
                            void* stream);

/* The same windows as candidate lists for the stateful searches: d_offsets[nq+1], and for query q the candidates
 * d_cands[d_offsets[q] .. d_offsets[q+1]) in GetFeaturesInArea order with their Hamming distances d_dist[]. Two passes
 * (count, scan, fill); synchronises the stream once to return the total (*total_out); ORB_ECAPACITY if total > cap
 * (call again with larger buffers). The caller then replays the reference's loop (e.g. src/ORBmatcher.cc:434-488). */
int orbm_window_lists_device(orbm_grid_handle g, const uint8_t* d_desc_frame, const uint8_t* d_queries, int nq, const float* d_x,
                             const float* d_y, const float* d_r, const int32_t* d_min_level, const int32_t* d_max_level,
                             int32_t* d_offsets, int32_t* d_cands, int16_t* d_dist, int cap, int32_t* total_out, void* stream);

/* Host-buffer form for callers whose frames live in host containers (the C++ facade on the reference's Frame): keypoints
 * (x, y, octave are read) with their descriptors, the image bounds {mnMinX, mnMinY, mnMaxX, mnMaxY} and nq windows go up in one
 * copy; grid build, window lists and distances run on the device; the lists come back in GetFeaturesInArea order. Per-thread
 * stream and buffers. ORB_ECAPACITY (with *total_out = needed room) when the lists do not fit `cap` entries. */
int orbm_window_lists(int device, const orbx_keypoint* kps, int n_kps, const float* bounds4, const uint8_t* desc_frame,
                      const uint8_t* queries, int nq, const float* x, const float* y, const float* r, const int32_t* min_level,
                      const int32_t* max_level, int32_t* offsets, int32_t* cands, int16_t* dist, int cap, int32_t* total_out);

/* ------------------------------------------------------------------------------------------
 * Map-point projection: Frame::isInFrustum (src/Frame.cc:269-325; mode 0) and the projection
 * prologue of ORBmatcher::Fuse / SearchByProjection(KF, Scw, ...) / SearchBySim3 (src/ORBmatcher.cc:849-889, 323-363,
 * 1169-1186; mode 1) for n map points kept as arrays on the device: world position and normal (n x 3 float),
 * mfMaxDistance / mfMinDistance (the raw members; the 1.2 / 0.8 factors of Get*DistanceInvariance are applied inside).
 * Outputs per point: alive (every gate the reference evaluates before GetFeaturesInArea passed), u, v, ur (= u - bf/z),
 * predicted level, viewing cosine; optionally the search window for orbm_window_*_device: radius (mode 0:
 * RadiusByViewingCos * th (th applied when != 1) * scale[level], src/ORBmatcher.cc:60-67,133-139; mode 1: th * scale[level]),
 * 0 when not alive, and the level window [level-1, level]. d_radius / d_min_level / d_max_level may be NULL. */
#define ORBM_MAX_LEVELS 32
typedef struct orbm_camera {
    float Rcw[9];   /* row-major rotation, world -> camera (Frame::mRcw / KeyFrame::GetRotation()) */
    float tcw[3];   /* translation (mtcw) */
    float Ow[3];    /* camera centre (mOw), as the caller's pose update computed it */
    float fx, fy, cx, cy, bf;
    float min_x, max_x, min_y, max_y; /* mnMinX ... mnMaxY */
    float log_scale_factor;            /* mfLogScaleFactor */
    int32_t n_levels;                  /* mnScaleLevels */
    float scale_factors[ORBM_MAX_LEVELS]; /* mvScaleFactors */
} orbm_camera;
int orbm_project_points_device(int device, const orbm_camera* cam, int mode, float viewing_cos_limit, float th,
                               const float* d_world_pos, const float* d_normal, const float* d_max_distance,
                               const float* d_min_distance, int n, uint8_t* d_alive, float* d_u, float* d_v, float* d_ur,
                               int32_t* d_level, float* d_view_cos, float* d_radius, int32_t* d_min_level,
                               int32_t* d_max_level, void* stream);

/* Frame::UndistortKeyPoints (src/Frame.cc:404-434): mvKeysUn for the keypoints an extractor left on the device (d_kps,
 * d_count, cap of orbx_device_results) - cv::undistortPoints(pts, pts, mK, mDistCoef, noArray(), mK) in OpenCV's double
 * arithmetic (5 iterations), x / y replaced, the other fields copied; a zero first coefficient is the reference's
 * "mvKeysUn = mvKeys" shortcut. dist_coef (HOST) = k1 k2 p1 p2 [k3 [k4 k5 k6]]. The result array is what
 * orbm_grid_build_device, the window searches and orbm_stereo_from_rgbd_device take for a distorted camera.
 * orbm_image_bounds (host only) is Frame::ComputeImageBounds (436-464): bounds4 = mnMinX, mnMaxX, mnMinY, mnMaxY. */
int orbm_undistort_keypoints_device(int device, const orbx_keypoint* d_kps, const int32_t* d_count, int cap, float fx, float fy, float cx,
                                    float cy, const float* dist_coef, int n_dist, orbx_keypoint* d_kps_un, void* stream);
int orbm_image_bounds(int width, int height, float fx, float fy, float cx, float cy, const float* dist_coef, int n_dist, float* bounds4);

/* Frame::ComputeStereoFromRGBD (src/Frame.cc:643-664) for the keypoints an extractor left on the device
 * (orbx_device_results: d_kps, d_count, cap): depth image (float, device, stride in bytes) sampled at the truncated RAW
 * keypoint position (mvKeys); d_depth[i] = d and d_uRight[i] = xU - mbf/d when d > 0, else -1 (also for i >= count), xU
 * taken from d_kps_un (mvKeysUn, orbm_undistort_keypoints_device) or from d_kps when d_kps_un is NULL (no distortion). */
int orbm_stereo_from_rgbd_device(int device, const orbx_keypoint* d_kps, const orbx_keypoint* d_kps_un, const int32_t* d_count, int cap,
                                 const float* d_depth_image, int width, int height, size_t stride_bytes, float mbf, float* d_uRight,
                                 float* d_depth, void* stream);

/* ------------------------------------------------------------------------------------------
 * Vocabulary tree descent (the step right after extraction): replaces the per-feature
 * TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup) with FORB::distance
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1259, FORB.cpp:81-101) for all features of a frame,
 * as called from Frame::ComputeBoW (src/Frame.cc:395-402). The vocabulary is given as the node table of the
 * ORBvoc text format (TemplatedVocabulary.h:1338-1424): nodes in id order, node 0 = root, parent[i] < i.
 * Outputs per feature: word id, node id at level L - levelsup, word weight. The caller folds them into its
 * DBoW2::BowVector / FeatureVector exactly like TemplatedVocabulary.h:1150-1163 does
 * (if (w > 0) { v.addWeight(id, w); fv.addFeature(nid, i); } ... v.normalize(norm)). */
typedef struct orbv_vocabulary* orbv_handle;
int orbv_create(int device, const int32_t* parent, const uint8_t* desc, const double* weight, int n_nodes, int k, int L,
                orbv_handle* out);
void orbv_destroy(orbv_handle v);
int orbv_descend_device(orbv_handle v, const uint8_t* d_desc, int n, int levelsup, int32_t* d_word, int32_t* d_node,
                        double* d_weight, void* stream);
int orbv_descend(orbv_handle v, const uint8_t* desc, int n, int levelsup, int32_t* word, int32_t* node, double* weight);

/* ------------------------------------------------------------------------------------------
 * Keyframe database scoring: for a query BowVector, the number of words shared with every
 * keyframe of the database (the inverted-file walk of KeyFrameDatabase::DetectLoopCandidates, src/KeyFrameDatabase.cc:85-105),
 * the first shared word (which fixes the order in which the reference meets the keyframes) and the similarity
 * mpVoc->score(query, keyframe) (131) = L1Scoring::score (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-66), bit-identical
 * doubles. BowVectors are given as (ascending word id, weight) arrays - the iteration order of DBoW2::BowVector.
 * orbdb_add replaces KeyFrameDatabase::add (40-46) and returns the keyframe's slot; orbdb_erase replaces erase (48-66).
 * The candidate bookkeeping that follows (minCommonWords, covisibility accumulation, 0.75*best filter) walks the
 * keyframe graph and stays with the caller. */
typedef struct orbdb_database* orbdb_handle;
int orbdb_create(int device, int n_words, orbdb_handle* out);
void orbdb_destroy(orbdb_handle db);
int orbdb_size(orbdb_handle db);
int orbdb_add(orbdb_handle db, const int32_t* word_ids, const double* weights, int n, int* slot_out);
int orbdb_erase(orbdb_handle db, int slot);
int orbdb_query_device(orbdb_handle db, const int32_t* q_word_ids, const double* q_weights, int nq, int32_t* d_common,
                       int32_t* d_first_word, double* d_score, void* stream);
int orbdb_query(orbdb_handle db, const int32_t* q_word_ids, const double* q_weights, int nq, int32_t* common, int32_t* first_word,
                double* score, int cap);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
