// TEST INFRASTRUCTURE ONLY (oracle). CPU restatement of the DBoW2 calls on the frontend path:
//   TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
//                                       /root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1196
//   TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)   :1218-1259
//   FORB::distance                      Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
//   BowVector::addWeight / normalize    Thirdparty/DBoW2/DBoW2/BowVector.cpp:34-82
//   FeatureVector::addFeature           Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:29-43
// as called by Frame::ComputeBoW (src/Frame.cc:395-402, levelsup = 4). TF-IDF weighting with L1
// normalisation (ORBvoc settings). Checked against the vendored DBoW2 itself (oracle/_ref/libref_dbow.so).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace {
struct Voc {
    int k, L, n;
    std::vector<int> parent, child_off, child_ids, word_id;
    std::vector<uint8_t> desc;
    std::vector<double> weight;
};
inline int ham(const uint8_t* a, const uint8_t* b) {
    uint32_t x[8], y[8];
    std::memcpy(x, a, 32); std::memcpy(y, b, 32);
    int d = 0;
    for (int i = 0; i < 8; ++i) d += __builtin_popcount(x[i] ^ y[i]);
    return d;
}
void descend(const Voc& v, const uint8_t* f, int levelsup, int& word, int& node, double& w) {
    const int nid_level = v.L - levelsup;
    node = 0;  // root when nid_level <= 0
    int cur = 0, level = 0;
    do {
        ++level;
        const int* ch = &v.child_ids[v.child_off[cur]];
        const int nch = v.child_off[cur + 1] - v.child_off[cur];
        cur = ch[0];
        double best = ham(f, &v.desc[(size_t)cur * 32]);
        for (int c = 1; c < nch; ++c) {
            const double d = ham(f, &v.desc[(size_t)ch[c] * 32]);
            if (d < best) { best = d; cur = ch[c]; }
        }
        if (level == nid_level) node = cur;
    } while (v.child_off[cur + 1] > v.child_off[cur]);
    word = v.word_id[cur];
    w = v.weight[cur];
}
}  // namespace

extern "C" {

// nodes in id order, node 0 = root; children lists are built in id order like loadFromTextFile does
void* orc_voc_create(const int32_t* parent, const uint8_t* desc, const double* weight, int n_nodes, int k, int L) {
    Voc* v = new Voc();
    v->k = k; v->L = L; v->n = n_nodes;
    v->parent.assign(parent, parent + n_nodes);
    v->desc.assign(desc, desc + (size_t)n_nodes * 32);
    v->weight.assign(weight, weight + n_nodes);
    std::vector<std::vector<int> > ch(n_nodes);
    for (int i = 1; i < n_nodes; ++i) ch[parent[i]].push_back(i);
    v->child_off.assign(n_nodes + 1, 0);
    for (int i = 0; i < n_nodes; ++i) { v->child_off[i + 1] = v->child_off[i] + (int)ch[i].size(); v->child_ids.insert(v->child_ids.end(), ch[i].begin(), ch[i].end()); }
    v->word_id.assign(n_nodes, 0);
    int w = 0;
    for (int i = 1; i < n_nodes; ++i) if (ch[i].empty()) v->word_id[i] = w++;  // words numbered in node order (1405-1412)
    return v;
}
void orc_voc_destroy(void* v) { delete (Voc*)v; }

void orc_voc_descend(void* vp, const uint8_t* desc, int n, int levelsup, int32_t* word, int32_t* node, double* weight) {
    const Voc& v = *(Voc*)vp;
    for (int i = 0; i < n; ++i) { int wd, nd; double w; descend(v, desc + (size_t)i * 32, levelsup, wd, nd, w); word[i] = wd; node[i] = nd; weight[i] = w; }
}

int orc_voc_transform(void* vp, const uint8_t* desc, int n, int levelsup, int32_t* bow_id, double* bow_val, int bow_cap, int* n_bow,
                      int32_t* fv_node, int32_t* fv_offset, int32_t* fv_feat, int fv_cap, int* n_fv) {
    const Voc& v = *(Voc*)vp;
    std::map<int, double> bow;
    std::map<int, std::vector<int> > fv;
    for (int i = 0; i < n; ++i) {
        int wd, nd; double w;
        descend(v, desc + (size_t)i * 32, levelsup, wd, nd, w);
        if (w > 0) { bow[wd] += w; fv[nd].push_back(i); }  // addWeight / addFeature
    }
    double norm = 0.0;  // L1 scoring: mustNormalize -> BowVector::normalize(L1)
    for (auto& e : bow) norm += std::fabs(e.second);
    if (norm > 0.0) for (auto& e : bow) e.second /= norm;
    int k = 0;
    for (auto& e : bow) { if (k < bow_cap) { bow_id[k] = e.first; bow_val[k] = e.second; } ++k; }
    *n_bow = k;
    int m = 0, pos = 0;
    fv_offset[0] = 0;
    for (auto& e : fv) {
        if (m < fv_cap) fv_node[m] = e.first;
        for (int f : e.second) fv_feat[pos++] = f;
        if (m < fv_cap) fv_offset[m + 1] = pos;
        ++m;
    }
    *n_fv = m;
    return 0;
}
}
