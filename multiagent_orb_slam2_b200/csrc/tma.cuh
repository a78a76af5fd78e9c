// TMA (cp.async.bulk.tensor, SASS UTMALDG) tile loads of u8 image tiles: 3-D tensor maps (x, y, frame)
// per pyramid level, zero fill outside the level. One elected thread issues the copy; completion is
// signalled on an mbarrier in shared memory.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace orb {

struct TmaMaps {
    CUtensorMap m[16];  // one per pyramid level (kMaxLevels)
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

// host: encode a (w, h, frames) u8 tensor with the given byte strides and box (bw x bh x 1)
int tma_encode_u8_3d(CUtensorMap* out, const void* base, int w, int h, int frames, size_t pitch, size_t frame_stride, int bw, int bh);

}  // namespace orb
