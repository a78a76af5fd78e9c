// TEST INFRASTRUCTURE ONLY (oracle). C entry points around the reference's vendored DBoW2
// (/root/reference/Thirdparty/DBoW2, compiled unmodified against oracle/shim by oracle/build_ref.sh):
// ORBVocabulary::loadFromTextFile + transform, i.e. what Frame::ComputeBoW (src/Frame.cc:395-402) calls.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/ScoringObject.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;  // include/ORBVocabulary.h

struct OpenVocabulary : public ORBVocabulary {  // exposes the protected per-feature descent
    void descend(const cv::Mat& f, DBoW2::WordId& id, DBoW2::WordValue& w, DBoW2::NodeId* nid, int levelsup) const {
        transform(f, id, w, nid, levelsup);
    }
};

extern "C" {

void* refv_load(const char* text_file) {
    OpenVocabulary* v = new OpenVocabulary();
    if (!v->loadFromTextFile(text_file)) { delete v; return nullptr; }
    return v;
}
void refv_destroy(void* v) { delete (OpenVocabulary*)v; }
int refv_size(void* v) { return (int)((OpenVocabulary*)v)->size(); }

static std::vector<cv::Mat> to_mats(const uint8_t* desc, int n) {
    std::vector<cv::Mat> f(n);
    for (int i = 0; i < n; ++i) { f[i].create(1, 32, CV_8U); std::memcpy(f[i].data, desc + (size_t)i * 32, 32); }
    return f;
}

// per-feature descent: word id, node id at level L - levelsup, word weight
void refv_descend(void* v, const uint8_t* desc, int n, int levelsup, int32_t* word, int32_t* node, double* weight) {
    const OpenVocabulary& voc = *(OpenVocabulary*)v;
    std::vector<cv::Mat> f = to_mats(desc, n);
    for (int i = 0; i < n; ++i) {
        DBoW2::WordId id; DBoW2::WordValue w; DBoW2::NodeId nid = 0;
        voc.descend(f[i], id, w, &nid, levelsup);
        word[i] = (int32_t)id; node[i] = (int32_t)nid; weight[i] = w;
    }
}

// the full transform: BowVector (ascending word id) and FeatureVector (ascending node id, CSR over feature indices)
int refv_transform(void* v, const uint8_t* desc, int n, int levelsup, int32_t* bow_id, double* bow_val, int bow_cap, int* n_bow,
                   int32_t* fv_node, int32_t* fv_offset, int32_t* fv_feat, int fv_cap, int* n_fv) {
    const OpenVocabulary& voc = *(OpenVocabulary*)v;
    std::vector<cv::Mat> f = to_mats(desc, n);
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    voc.ORBVocabulary::transform(f, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k)
        if (k < bow_cap) { bow_id[k] = (int32_t)it->first; bow_val[k] = it->second; }
    *n_bow = k;
    int m = 0, pos = 0;
    fv_offset[0] = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++m) {
        if (m < fv_cap) fv_node[m] = (int32_t)it->first;
        for (size_t j = 0; j < it->second.size(); ++j, ++pos) fv_feat[pos] = (int32_t)it->second[j];
        if (m < fv_cap) fv_offset[m + 1] = pos;
    }
    *n_fv = m;
    return 0;
}

// mpVoc->score(v1, v2) for the L1_NORM scoring of the ORB vocabulary: the vendored L1Scoring::score on two BowVectors
double refv_l1_score(const int32_t* id1, const double* w1, int n1, const int32_t* id2, const double* w2, int n2) {
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; ++i) a.addWeight((DBoW2::WordId)id1[i], w1[i]);
    for (int i = 0; i < n2; ++i) b.addWeight((DBoW2::WordId)id2[i], w2[i]);
    return DBoW2::L1Scoring().score(a, b);
}

}  // extern "C"
