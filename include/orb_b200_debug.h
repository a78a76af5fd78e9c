/* orb_b200_debug.h - test taps and measurement hooks of liborb_b200.so.
 *
 * NOT part of the drop-in boundary (include/orb_b200.h): nothing here replaces a reference interface, a SLAM caller never
 * needs it. The parity tests read intermediate stages through it, bench.py reads per-stage device times, and the
 * measurements pin one of the two brute-force Hamming implementations. Same conventions as orb_b200.h. */
#ifndef ORB_B200_DEBUG_H
#define ORB_B200_DEBUG_H

#include "orb_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Stage taps used by the parity tests (not needed by a SLAM caller): the blurred level
 * (src/ORBextractor.cc:1085-1086) and the FAST candidates of a level in the reference's order
 * (cell row-major, then row-major inside the cell; x,y in level coordinates; 789-829). */
int orbx_debug_blurred_level(orbx_handle h, int frame, int level, uint8_t* dst, size_t dst_stride);
int orbx_debug_candidates(orbx_handle h, int frame, int level, int32_t* xys /* [cap][3] */, int cap, int* n_out);

/* Measurement hooks (bench.py): with stage timing on, the pipeline runs its stages back to back on
 * ONE stream with CUDA events in between; orbx_stage_times returns the device milliseconds of
 * {resize chain, blur, FAST, quadtree, orient+describe} of the last call, orbx_algorithmic_bytes the
 * per-frame algorithmic HBM bytes of the same stages. */
#define ORBX_NUM_STAGES 5
int orbx_set_stage_timing(orbx_handle h, int enable);
int orbx_stage_times(orbx_handle h, float* ms /* [ORBX_NUM_STAGES] */);
int orbx_algorithmic_bytes(orbx_handle h, double* bytes /* [ORBX_NUM_STAGES] */);

/* Stand-alone quadtree stage on host buffers (parity tests / micro-benchmarks). */
int orbx_debug_quadtree(int device, const int32_t* xs, const int32_t* ys, const int32_t* scores, int n,
                        int minX, int maxX, int minY, int maxY, int N, int32_t* out_idx, int cap, int* n_out);

/* The brute-force searches (orbm_knn2*, orbm_knn2_pairs_device ...) have two implementations with identical results: the
 * POPC kernel (csrc/hamming.cu) and the tensor-core kernel (csrc/hamming_mma.cu), chosen by problem size. This pins the
 * choice FOR THE CALLING THREAD ONLY (thread-local; other matcher threads keep the automatic choice):
 * 0 = automatic (default), 1 = POPC only, 2 = tensor cores with one CTA per 128 query rows (tcgen05 cta_group::1),
 * 3 = tensor cores with a CTA pair per 256 query rows (cta_group::2; automatic choice for long databases). */
int orbm_set_knn2_backend(int backend);

/* Occupancy of the two tensor-core matcher kernels on the current device: resident CTA pairs (clusters) on the whole GPU and
 * resident blocks per SM. */
int orbm_debug_mma_occupancy(int* pair_clusters, int* single_blocks_per_sm);
/* The tensor-core implementation called directly (one query set against one database set), whatever the size. */
int orbm_knn2_mma_device(const uint8_t* dA, int nA, const uint8_t* dB, int nB, int32_t* d_idx, int32_t* d_best,
                         int32_t* d_second, void* stream);
/* Debug tap of its data path: the +-1 dot products (= 256 - 2 * distance) of 128 x 256 host descriptors through the bit
 * expansion, the SWIZZLE_128B TMA loads, tcgen05.mma kind::i8 and tcgen05.ld; out = int32 [128][256]. */
int orbm_debug_mma_dot(int device, const uint8_t* A128, const uint8_t* B256, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_DEBUG_H */
