// C++ driver of the cross-map entry points of include/orb_b200.h (orbm_xmap_* / orbm_knn2_allgather), the way a single-process
// multi-agent server (the reference's MultiAgentServer + MapFusion thread, src/MapFusion.cc:275, 849) would use them: one
// context per visible GPU attached with orbm_xmap_attach_local, n_maps maps (fewer than GPUs is allowed: query rows are
// split), every pair checked on sampled rows against a scalar popcount loop.
//   xmap_test <n_maps> <rows> [max_gpus]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "orb_b200.h"

#define CHECK(x) do { int _rc = (x); if (_rc != ORB_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, _rc, orb_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(_e)); return 1; } } while (0)

static int hamming(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount(a[i] ^ b[i]);
    return d;
}

int main(int argc, char** argv) {
    const int n_maps = argc > 1 ? atoi(argv[1]) : 2, rows = argc > 2 ? atoi(argv[2]) : 5000;
    int world = orb_device_count();
    if (argc > 3 && atoi(argv[3]) < world) world = atoi(argv[3]);
    if (world < 1) { fprintf(stderr, "no CUDA device\n"); return 1; }
    std::mt19937 rng(7);
    std::vector<std::vector<uint8_t> > maps(n_maps);
    std::vector<int32_t> rpm(n_maps);
    for (int m = 0; m < n_maps; ++m) {
        rpm[m] = rows - 11 * m;
        maps[m].resize((size_t)rpm[m] * 32);
        for (size_t i = 0; i < maps[m].size(); ++i) maps[m][i] = (uint8_t)rng();
        if (m) for (int r = 0; r < rpm[m]; r += 3) {   // near-duplicates of map 0 rows so that best << second for many rows
            memcpy(&maps[m][(size_t)r * 32], &maps[0][(size_t)(r % rpm[0]) * 32], 32);
            maps[m][(size_t)r * 32 + (r % 32)] ^= (uint8_t)(1u << (r % 8));
        }
    }
    std::vector<orbm_xmap_t> ctx(world);
    for (int r = 0; r < world; ++r) CHECK(orbm_xmap_create(r, r, world, n_maps, rows, &ctx[r]));
    CHECK(orbm_xmap_attach_local(ctx.data(), world));
    std::vector<uint8_t*> d(n_maps);
    std::vector<cudaStream_t> st(world);
    for (int r = 0; r < world; ++r) { CU(cudaSetDevice(r)); CU(cudaStreamCreateWithFlags(&st[r], cudaStreamNonBlocking)); }
    for (int m = 0; m < n_maps; ++m) {
        CU(cudaSetDevice(m % world));
        CU(cudaMalloc(&d[m], maps[m].size()));
        CU(cudaMemcpy(d[m], maps[m].data(), maps[m].size(), cudaMemcpyHostToDevice));
    }
    for (int step = 0; step < 3; ++step) {
        for (int r = 0; r < world; ++r) {
            std::vector<const uint8_t*> own;
            for (int m = r; m < n_maps; m += world) own.push_back(d[m]);
            CHECK(orbm_knn2_allgather(ctx[r], own.data(), rpm.data(), st[r]));
        }
        for (int r = 0; r < world; ++r) { CU(cudaSetDevice(r)); CU(cudaStreamSynchronize(st[r])); }
    }
    long checked = 0;
    for (int a = 0; a < n_maps; ++a)
        for (int b = 0; b < n_maps; ++b) {
            if (a == b) continue;
            const int32_t *di, *d1, *d2;
            CHECK(orbm_xmap_result(ctx[a % world], a, b, &di, &d1, &d2));
            std::vector<int32_t> hi(rpm[a]), h1(rpm[a]), h2(rpm[a]);
            CU(cudaSetDevice(a % world));
            CU(cudaMemcpy(hi.data(), di, hi.size() * 4, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(h2.data(), d2, h2.size() * 4, cudaMemcpyDeviceToHost));
            for (int q = 0; q < rpm[a]; q += 97) {   // reference selection rule: strict '<', first minimum, both start at 256
                int best = 256, second = 256, idx = -1;
                for (int j = 0; j < rpm[b]; ++j) {
                    const int dist = hamming(&maps[a][(size_t)q * 32], &maps[b][(size_t)j * 32]);
                    if (dist < best) { second = best; best = dist; idx = j; }
                    else if (dist < second) second = dist;
                }
                if (hi[q] != idx || h1[q] != best || h2[q] != second) {
                    fprintf(stderr, "pair (%d,%d) row %d: got (%d,%d,%d) want (%d,%d,%d)\n", a, b, q, hi[q], h1[q], h2[q], idx, best, second);
                    return 1;
                }
                ++checked;
            }
        }
    for (int r = 0; r < world; ++r) orbm_xmap_destroy(ctx[r]);
    printf("xmap ok: %d maps x %d rows on %d GPU(s), %ld sampled rows identical\n", n_maps, rows, world, checked);
    return 0;
}
