// TEST INFRASTRUCTURE ONLY (oracle). Multi-threaded drivers around the scalar oracle loops,
// used for the CPU baseline legs of bench.py (rows of the query set / frames of a batch are split
// across host threads; each thread runs the unmodified scalar restatement).
#include <cstdint>
#include <cstddef>
#include <thread>
#include <vector>

extern "C" {
void orc_knn2(const uint8_t* A, int nA, const uint8_t* B, int nB, int* idx, int* d1, int* d2);
void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const int32_t* pattern1024);
void orc_extractor_destroy(void* h);
int orc_extract(void* h, const uint8_t* img, int w, int ht, size_t stride);

void orc_knn2_mt(const uint8_t* A, int nA, const uint8_t* B, int nB, int* idx, int* d1, int* d2, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    const int per = (nA + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int lo = t * per, hi = lo + per < nA ? lo + per : nA;
        if (lo >= hi) break;
        pool.emplace_back([=] { orc_knn2(A + (size_t)lo * 32, hi - lo, B, nB, idx + lo, d1 + lo, d2 + lo); });
    }
    for (auto& th : pool) th.join();
}

// extracts `n` frames (contiguous, w*h each) with one extractor per thread; returns total keypoints
long orc_extract_mt(const uint8_t* imgs, int n, int w, int ht, int nfeatures, float scaleFactor, int nlevels,
                    int iniTh, int minTh, const int32_t* pattern1024, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<long> total(nthreads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([=, &total] {
            void* e = orc_extractor_create(nfeatures, scaleFactor, nlevels, iniTh, minTh, pattern1024);
            for (int i = t; i < n; i += nthreads) total[t] += orc_extract(e, imgs + (size_t)i * w * ht, w, ht, w);
            orc_extractor_destroy(e);
        });
    }
    for (auto& th : pool) th.join();
    long s = 0;
    for (long v : total) s += v;
    return s;
}
}
