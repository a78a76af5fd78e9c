// Keyframe database scoring: the word-sharing count and the DBoW2 L1 similarity score of a query
// BowVector against every keyframe of the database in one launch - the arithmetic inside
// KeyFrameDatabase::DetectLoopCandidates / DetectCovisibilityCandidates / DetectRelocalizationCandidates
// (/root/reference/src/KeyFrameDatabase.cc:76-197, 199-308, 310-420): the inverted-file walk that counts common words
// (85-105) and mpVoc->score(pKF->mBowVec, pKFi->mBowVec) (131) = L1Scoring::score
// (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-66).
//
// Layout: the keyframes' BowVectors as CSR (word ids ascending, double weights) in HBM; the query scattered into a dense
// table of n_words doubles (8 MB for the 10^6-word ORB vocabulary, L2 resident). One warp per keyframe: lanes read 32
// consecutive (word, weight) entries (coalesced), look their word up in the table, and the terms of the common words are
// added in ascending word order - the order of the reference's merge walk - so the double sum is bit-identical.
#include <vector>

#include <cstring>

#include "common.cuh"

struct orbdb_database {
    int device = 0;
    int n_words = 0;
    std::vector<long long> offsets;  // host CSR, n_kf + 1
    std::vector<int32_t> words;
    std::vector<double> weights;
    std::vector<uint8_t> alive;
    // device mirror
    long long* d_offsets = nullptr; int32_t* d_words = nullptr; double* d_weights = nullptr; uint8_t* d_alive = nullptr;
    size_t cap_kf = 0, cap_entries = 0, uploaded_kf = 0, uploaded_entries = 0;
    bool alive_dirty = false;
    double* d_table = nullptr;   // dense query weights, zero outside a query
    int32_t* d_qwords = nullptr; double* d_qweights = nullptr; size_t cap_q = 0;
};

namespace orb {

__global__ void kfdb_scatter_kernel(double* __restrict__ table, const int32_t* __restrict__ qw, const double* __restrict__ qv, int nq, int clear) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) table[qw[i]] = clear ? 0.0 : qv[i];
}

__global__ void __launch_bounds__(256) kfdb_score_kernel(const long long* __restrict__ offsets, const int32_t* __restrict__ words,
                                                         const double* __restrict__ weights, const uint8_t* __restrict__ alive, int n_kf,
                                                         const double* __restrict__ table, int32_t* __restrict__ common,
                                                         int32_t* __restrict__ first_word, double* __restrict__ score) {
    const int kf = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (kf >= n_kf) return;
    if (!alive[kf]) {
        if (lane == 0) { common[kf] = 0; first_word[kf] = -1; score[kf] = 0.0; }
        return;
    }
    const long long b = offsets[kf], e = offsets[kf + 1];
    double s = 0.0;
    int n_common = 0, first = -1;
    for (long long base = b; base < e; base += 32) {
        const long long i = base + lane;
        double term = 0.0;
        int w = -1;
        bool hit = false;
        if (i < e) {
            w = words[i];
            const double vi = table[w];  // the query's weight (v1), 0 when the query does not have the word
            if (vi != 0.0) {
                const double wi = weights[i];
                term = __dsub_rn(__dsub_rn(fabs(__dsub_rn(vi, wi)), fabs(vi)), fabs(wi));
                hit = true;
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m && first < 0) first = __shfl_sync(0xffffffffu, w, __ffs(m) - 1);
        n_common += __popc(m);
        while (m) {  // ascending word order
            const int src = __ffs(m) - 1;
            m &= m - 1;
            s = __dadd_rn(s, __shfl_sync(0xffffffffu, term, src));
        }
    }
    if (lane == 0) {
        common[kf] = n_common;
        first_word[kf] = first;
        score[kf] = -s / 2.0;
    }
}

static int kfdb_sync(orbdb_database* db, cudaStream_t st) {
    const size_t n_kf = db->alive.size(), n_ent = db->words.size();
    if (n_kf + 1 > db->cap_kf || n_ent > db->cap_entries) {  // grow geometrically, re-upload everything
        const size_t ck = std::max<size_t>(1024, 2 * (n_kf + 1)), ce = std::max<size_t>(1 << 16, 2 * n_ent);
        long long* o; int32_t* w; double* v; uint8_t* a;
        ORB_CUDA_TRY(cudaMalloc(&o, ck * sizeof(long long)));
        ORB_CUDA_TRY(cudaMalloc(&w, ce * sizeof(int32_t)));
        ORB_CUDA_TRY(cudaMalloc(&v, ce * sizeof(double)));
        ORB_CUDA_TRY(cudaMalloc(&a, ck));
        cudaFree(db->d_offsets); cudaFree(db->d_words); cudaFree(db->d_weights); cudaFree(db->d_alive);
        db->d_offsets = o; db->d_words = w; db->d_weights = v; db->d_alive = a;
        db->cap_kf = ck; db->cap_entries = ce; db->uploaded_kf = 0; db->uploaded_entries = 0; db->alive_dirty = true;
    }
    if (db->uploaded_kf < n_kf) {
        ORB_CUDA_TRY(cudaMemcpyAsync(db->d_offsets + db->uploaded_kf, db->offsets.data() + db->uploaded_kf,
                                     (n_kf + 1 - db->uploaded_kf) * sizeof(long long), cudaMemcpyHostToDevice, st));
        ORB_CUDA_TRY(cudaMemcpyAsync(db->d_words + db->uploaded_entries, db->words.data() + db->uploaded_entries,
                                     (n_ent - db->uploaded_entries) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        ORB_CUDA_TRY(cudaMemcpyAsync(db->d_weights + db->uploaded_entries, db->weights.data() + db->uploaded_entries,
                                     (n_ent - db->uploaded_entries) * sizeof(double), cudaMemcpyHostToDevice, st));
        db->uploaded_kf = n_kf; db->uploaded_entries = n_ent; db->alive_dirty = true;
    }
    if (db->alive_dirty && n_kf) {
        ORB_CUDA_TRY(cudaMemcpyAsync(db->d_alive, db->alive.data(), n_kf, cudaMemcpyHostToDevice, st));
        db->alive_dirty = false;
    }
    return ORB_OK;
}

}  // namespace orb

extern "C" {

int orbdb_create(int device, int n_words, orbdb_handle* out) {
    using namespace orb;
    ORB_REQUIRE(out && n_words > 0, "bad arguments");
    ORB_CUDA_TRY(cudaSetDevice(device));
    orbdb_database* db = new orbdb_database();
    db->device = device; db->n_words = n_words;
    db->offsets.push_back(0);
    if (cudaMalloc(&db->d_table, (size_t)n_words * sizeof(double)) != cudaSuccess ||
        cudaMemset(db->d_table, 0, (size_t)n_words * sizeof(double)) != cudaSuccess) {
        set_error("orbdb_create: cannot allocate the %d-word query table", n_words);
        delete db;
        return ORB_ECUDA;
    }
    *out = db;
    return ORB_OK;
}

void orbdb_destroy(orbdb_handle db) {
    if (!db) return;
    cudaSetDevice(db->device);
    cudaFree(db->d_offsets); cudaFree(db->d_words); cudaFree(db->d_weights); cudaFree(db->d_alive);
    cudaFree(db->d_table); cudaFree(db->d_qwords); cudaFree(db->d_qweights);
    delete db;
}

int orbdb_size(orbdb_handle db) { return db ? (int)db->alive.size() : 0; }

int orbdb_add(orbdb_handle db, const int32_t* word_ids, const double* weights, int n, int* slot_out) {
    using namespace orb;
    ORB_REQUIRE(db && n >= 0 && (n == 0 || (word_ids && weights)), "bad arguments");
    for (int i = 0; i < n; ++i) {
        ORB_REQUIRE(word_ids[i] >= 0 && word_ids[i] < db->n_words, "word id outside the vocabulary");
        ORB_REQUIRE(i == 0 || word_ids[i] > word_ids[i - 1], "BowVector must be sorted by word id (std::map order)");
    }
    db->words.insert(db->words.end(), word_ids, word_ids + n);
    db->weights.insert(db->weights.end(), weights, weights + n);
    db->offsets.push_back((long long)db->words.size());
    db->alive.push_back(1);
    if (slot_out) *slot_out = (int)db->alive.size() - 1;
    return ORB_OK;
}

int orbdb_erase(orbdb_handle db, int slot) {
    using namespace orb;
    ORB_REQUIRE(db && slot >= 0 && slot < (int)db->alive.size(), "bad slot");
    db->alive[slot] = 0;
    db->alive_dirty = true;
    return ORB_OK;
}

int orbdb_query_device(orbdb_handle db, const int32_t* q_word_ids, const double* q_weights, int nq, int32_t* d_common,
                       int32_t* d_first_word, double* d_score, void* stream) {
    using namespace orb;
    ORB_REQUIRE(db && nq >= 0 && (nq == 0 || (q_word_ids && q_weights)) && d_common && d_first_word && d_score, "bad arguments");
    for (int i = 0; i < nq; ++i) ORB_REQUIRE(q_word_ids[i] >= 0 && q_word_ids[i] < db->n_words, "word id outside the vocabulary");
    ORB_CUDA_TRY(cudaSetDevice(db->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = kfdb_sync(db, st);
    if (rc != ORB_OK) return rc;
    const int n_kf = (int)db->alive.size();
    if (n_kf == 0) return ORB_OK;
    if ((size_t)nq > db->cap_q) {
        cudaFree(db->d_qwords); cudaFree(db->d_qweights);
        db->cap_q = std::max<size_t>(4096, 2 * (size_t)nq);
        ORB_CUDA_TRY(cudaMalloc(&db->d_qwords, db->cap_q * sizeof(int32_t)));
        ORB_CUDA_TRY(cudaMalloc(&db->d_qweights, db->cap_q * sizeof(double)));
    }
    if (nq) {
        ORB_CUDA_TRY(cudaMemcpyAsync(db->d_qwords, q_word_ids, (size_t)nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        ORB_CUDA_TRY(cudaMemcpyAsync(db->d_qweights, q_weights, (size_t)nq * sizeof(double), cudaMemcpyHostToDevice, st));
        kfdb_scatter_kernel<<<ceil_div(nq, 256), 256, 0, st>>>(db->d_table, db->d_qwords, db->d_qweights, nq, 0);
    }
    kfdb_score_kernel<<<ceil_div(n_kf, 8), 256, 0, st>>>(db->d_offsets, db->d_words, db->d_weights, db->d_alive, n_kf, db->d_table, d_common,
                                                        d_first_word, d_score);
    if (nq) kfdb_scatter_kernel<<<ceil_div(nq, 256), 256, 0, st>>>(db->d_table, db->d_qwords, db->d_qweights, nq, 1);
    count_launch(nq ? 3 : 1);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbdb_query(orbdb_handle db, const int32_t* q_word_ids, const double* q_weights, int nq, int32_t* common, int32_t* first_word,
                double* score, int cap) {
    using namespace orb;
    ORB_REQUIRE(db && common && first_word && score, "null pointer");
    const int n_kf = (int)db->alive.size();
    ORB_REQUIRE(cap >= n_kf, "output arrays smaller than the database");
    if (n_kf == 0) return ORB_OK;
    ORB_CUDA_TRY(cudaSetDevice(db->device));
    // scores | common words | first word: the calling thread's arena (common.cuh), one copy back through pinned staging
    HostCallWorkspace& ws = host_call_workspace();
    int rc;
    if ((rc = ws.begin(db->device, ws.need((size_t)n_kf * 16)))) return rc;
    uint8_t* d = ws.take<uint8_t>((size_t)n_kf * 16);
    double* d_score = reinterpret_cast<double*>(d);
    int32_t* d_common = reinterpret_cast<int32_t*>(d_score + n_kf);
    int32_t* d_first = d_common + n_kf;
    if ((rc = orbdb_query_device(db, q_word_ids, q_weights, nq, d_common, d_first, d_score, nullptr))) return rc;   // default stream
    uint8_t* p = nullptr;
    if ((rc = ws.pinned((size_t)n_kf * 16, &p))) return rc;
    ORB_CUDA_TRY(cudaMemcpyAsync(p, d, (size_t)n_kf * 16, cudaMemcpyDeviceToHost, nullptr));
    ORB_CUDA_TRY(cudaStreamSynchronize(nullptr));
    memcpy(score, p, (size_t)n_kf * 8);
    memcpy(common, p + (size_t)n_kf * 8, (size_t)n_kf * 4);
    memcpy(first_word, p + (size_t)n_kf * 12, (size_t)n_kf * 4);
    return ORB_OK;
}

}  // extern "C"
