// K10: DBoW2 vocabulary-tree descent of ORB descriptors (the Hamming-bound part of Frame::ComputeBoW).
//
// Replaces TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)
// (/root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1259) with FORB::distance
// (Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101) for every feature of a frame, as called through
// transform(features, BowVector&, FeatureVector&, levelsup) (1127-1196) from Frame::ComputeBoW
// (src/Frame.cc:395-402). Per level the feature is compared with the <= k children of the current
// node and moves to the first child of minimum distance (strict '<' update, 1240-1248).
// A group of G lanes (G = 16 for k = 10) serves one feature: lane c computes the distance to child c,
// the group reduces (distance << 8 | c) with a minimum. Integer only: bit-exact.
// The BowVector / FeatureVector assembly (addWeight, addFeature, L1 normalisation - a few hundred
// double additions in a fixed order) stays on the host, where the caller's DBoW2 containers live.
#include <new>
#include <vector>

#include <cstring>

#include "common.cuh"

struct orbv_vocabulary {
    int device = 0, k = 0, L = 0, n_nodes = 0, group = 32;
    uint8_t* d_desc = nullptr;
    int* d_child_off = nullptr; int* d_child_ids = nullptr; int* d_word = nullptr;
    double* d_weight = nullptr;
};

namespace orb {

template <int G>
__global__ void __launch_bounds__(256)
bow_descend_kernel(const uint4* __restrict__ feats, int n, const uint4* __restrict__ node_desc, const int* __restrict__ child_off,
                   const int* __restrict__ child_ids, const int* __restrict__ word_id, const double* __restrict__ weight, int nid_level,
                   int* __restrict__ o_word, int* __restrict__ o_node, double* __restrict__ o_weight) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = t / G, c = t % G;
    const bool live = f < n;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (live) { q0 = __ldg(feats + (size_t)f * 2); q1 = __ldg(feats + (size_t)f * 2 + 1); }
    int cur = 0, level = 0, node_out = 0;
    for (;;) {
        const int off = live ? child_off[cur] : 0, nch = live ? child_off[cur + 1] - off : 0;
        if (__all_sync(0xffffffffu, nch == 0)) break;  // every feature of the warp sits on a leaf
        uint32_t key = 0xffffffffu;
        if (c < nch) {
            const int id = child_ids[off + c];
            const uint4 b0 = __ldg(node_desc + (size_t)id * 2), b1 = __ldg(node_desc + (size_t)id * 2 + 1);
            const uint32_t d = __popc(q0.x ^ b0.x) + __popc(q0.y ^ b0.y) + __popc(q0.z ^ b0.z) + __popc(q0.w ^ b0.w) +
                               __popc(q1.x ^ b1.x) + __popc(q1.y ^ b1.y) + __popc(q1.z ^ b1.z) + __popc(q1.w ^ b1.w);
            key = d << 8 | (uint32_t)c;
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
        if (nch > 0) {
            ++level;
            cur = child_ids[off + (int)(key & 0xffu)];
            if (level == nid_level) node_out = cur;
        }
    }
    if (live && c == 0) { o_word[f] = word_id[cur]; o_node[f] = node_out; o_weight[f] = weight[cur]; }
}

}  // namespace orb

using namespace orb;

extern "C" {

int orbv_create(int device, const int32_t* parent, const uint8_t* desc, const double* weight, int n_nodes, int k, int L, orbv_handle* out) {
    ORB_REQUIRE(parent && desc && weight && out && n_nodes >= 2 && k >= 1 && L >= 1, "bad vocabulary arguments");
    *out = nullptr;
    // children lists in node id order, words numbered in node id order: exactly what loadFromTextFile builds (1385-1412)
    std::vector<std::vector<int> > ch(n_nodes);
    for (int i = 1; i < n_nodes; ++i) {
        ORB_REQUIRE(parent[i] >= 0 && parent[i] < i, "parents must precede their children (node id order of the text format)");
        ch[parent[i]].push_back(i);
    }
    std::vector<int> off(n_nodes + 1, 0), ids, word(n_nodes, 0);
    int maxch = 0, w = 0;
    for (int i = 0; i < n_nodes; ++i) {
        off[i + 1] = off[i] + (int)ch[i].size();
        ids.insert(ids.end(), ch[i].begin(), ch[i].end());
        maxch = std::max(maxch, (int)ch[i].size());
        if (i > 0 && ch[i].empty()) word[i] = w++;
    }
    ORB_REQUIRE(maxch >= 1 && maxch <= 32, "a node has more than 32 children");
    ORB_CUDA_TRY(cudaSetDevice(device));
    orbv_vocabulary* v = new (std::nothrow) orbv_vocabulary();
    ORB_REQUIRE(v, "out of host memory");
    v->device = device; v->k = k; v->L = L; v->n_nodes = n_nodes;
    v->group = maxch <= 8 ? 8 : (maxch <= 16 ? 16 : 32);
    cudaError_t e = cudaMalloc(&v->d_desc, (size_t)n_nodes * 32);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_child_off, (size_t)(n_nodes + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_child_ids, std::max<size_t>(ids.size(), 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_word, (size_t)n_nodes * 4);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_weight, (size_t)n_nodes * 8);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_desc, desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_word, word.data(), word.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_weight, weight, (size_t)n_nodes * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("vocabulary upload failed: %s", cudaGetErrorString(e)); orbv_destroy(v); return ORB_ECUDA; }
    *out = v;
    return ORB_OK;
}

void orbv_destroy(orbv_handle v) {
    if (!v) return;
    cudaSetDevice(v->device);
    void* p[] = {v->d_desc, v->d_child_off, v->d_child_ids, v->d_word, v->d_weight};
    for (void* q : p) if (q) cudaFree(q);
    delete v;
}

int orbv_descend_device(orbv_handle v, const uint8_t* d_desc, int n, int levelsup, int32_t* d_word, int32_t* d_node, double* d_weight,
                        void* stream) {
    ORB_REQUIRE(v && n >= 0, "bad arguments");
    if (n == 0) return ORB_OK;
    ORB_REQUIRE(d_desc && d_word && d_node && d_weight, "null pointer");
    ORB_CUDA_TRY(cudaSetDevice(v->device));
    const int nid_level = v->L - levelsup;  // level whose node id goes into the FeatureVector (<= 0: root)
    cudaStream_t st = (cudaStream_t)stream;
    const int G = v->group;
    const int blocks = ceil_div(n * G, 256);
#define ORBV_LAUNCH(GG)                                                                                                              \
    bow_descend_kernel<GG><<<blocks, 256, 0, st>>>((const uint4*)d_desc, n, (const uint4*)v->d_desc, v->d_child_off, v->d_child_ids, \
                                                   v->d_word, v->d_weight, nid_level, d_word, d_node, d_weight)
    if (G == 8) ORBV_LAUNCH(8); else if (G == 16) ORBV_LAUNCH(16); else ORBV_LAUNCH(32);
#undef ORBV_LAUNCH
    count_launch();
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbv_descend(orbv_handle v, const uint8_t* desc, int n, int levelsup, int32_t* word, int32_t* node, double* weight) {
    ORB_REQUIRE(v && n >= 0, "bad arguments");
    if (n == 0) return ORB_OK;
    ORB_REQUIRE(desc && word && node && weight, "null pointer");
    // the vocabulary is shared by every thread of a SLAM system (ComputeBoW): the calling thread's own arena, stream and pinned
    // staging (common.cuh) - descriptors up, one copy of weight | word | node back, one synchronisation
    HostCallWorkspace& ws = host_call_workspace();
    int rc;
    if ((rc = ws.begin(v->device, ws.need((size_t)n * 32) + ws.need((size_t)n * 16)))) return rc;
    uint8_t* d_desc = ws.take<uint8_t>((size_t)n * 32);
    uint8_t* d_out = ws.take<uint8_t>((size_t)n * 16);
    double* dwt = reinterpret_cast<double*>(d_out);
    int32_t* dw = reinterpret_cast<int32_t*>(d_out + (size_t)n * 8);
    ORB_CUDA_TRY(cudaMemcpyAsync(d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, ws.st));
    if ((rc = orbv_descend_device(v, d_desc, n, levelsup, dw, dw + n, dwt, ws.st))) return rc;
    uint8_t* p = nullptr;
    if ((rc = ws.pinned((size_t)n * 16, &p))) return rc;
    ORB_CUDA_TRY(cudaMemcpyAsync(p, d_out, (size_t)n * 16, cudaMemcpyDeviceToHost, ws.st));
    ORB_CUDA_TRY(cudaStreamSynchronize(ws.st));
    memcpy(weight, p, (size_t)n * 8);
    memcpy(word, p + (size_t)n * 8, (size_t)n * 4);
    memcpy(node, p + (size_t)n * 12, (size_t)n * 4);
    return ORB_OK;
}

}  // extern "C"
