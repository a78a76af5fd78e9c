// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// CPU restatement of the reference's ORB frontend hot path, written against raw arrays:
//   ORBextractor ctor tables       /root/reference/src/ORBextractor.cc:410-470
//   ComputePyramid                 src/ORBextractor.cc:1107-1132
//   ComputeKeyPointsOctTree        src/ORBextractor.cc:765-853
//   DistributeOctTree / DivideNode src/ORBextractor.cc:481-763
//   IC_Angle                       src/ORBextractor.cc:77-104
//   computeOrbDescriptor           src/ORBextractor.cc:108-147
//   operator()                     src/ORBextractor.cc:1043-1105
//   ORBmatcher::DescriptorDistance src/ORBmatcher.cc:1649-1665
// OpenCV primitives come from cvprim.h (pinned to cv2 4.13.0).
//
// Parity status: the reference ships no tests / golden vectors, so the pin
// is (i) cv2 4.13.0 for the OpenCV primitives and (ii) the reference's own ORBextractor.cc
// compiled unmodified against oracle/shim (oracle/_ref, see build_ref.sh) for everything else.
//
// Documented deviations (canonical forms of behaviour the reference leaves unspecified):
//   * DistributeOctTree sorts pair<int,ExtractorNode*> (src/ORBextractor.cc:684): ties in node
//     size fall to heap addresses. Canonical rule here: stable sort on size only (ties keep
//     creation order). oracle/_ref builds both the verbatim and the canonical variant.
//   * float expressions are evaluated without FMA contraction (build with -ffp-contract=off).
//
// The BRIEF sampling pattern (src/ORBextractor.cc:150-408) is data, loaded from
// oracle/orb_pattern.bin at run time (1024 little-endian int32).
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

#include "cvprim.h"

namespace {

constexpr int kPatch = 31;
constexpr int kHalfPatch = 15;
constexpr int kEdge = 19;
constexpr int kCell = 30;

struct Cand { int x, y, score; };  // x,y relative to (minBorderX,minBorderY) like the reference

// ------------------------------------------------------------------------------------------
// Quadtree keypoint distribution (src/ORBextractor.cc:481-763), canonical tie-break.
struct QNode {
    int x0, x1, y0, y1;
    std::vector<int> members;  // candidate indices, candidate order preserved
    bool frozen = false;
    std::list<QNode>::iterator self;
};

void split_node(const QNode& n, const std::vector<Cand>& c, QNode ch[4]) {
    const int hx = (int)std::ceil((float)(n.x1 - n.x0) / 2);
    const int hy = (int)std::ceil((float)(n.y1 - n.y0) / 2);
    const int xm = n.x0 + hx, ym = n.y0 + hy;
    ch[0].x0 = n.x0; ch[0].x1 = xm;   ch[0].y0 = n.y0; ch[0].y1 = ym;
    ch[1].x0 = xm;   ch[1].x1 = n.x1; ch[1].y0 = n.y0; ch[1].y1 = ym;
    ch[2].x0 = n.x0; ch[2].x1 = xm;   ch[2].y0 = ym;   ch[2].y1 = n.y1;
    ch[3].x0 = xm;   ch[3].x1 = n.x1; ch[3].y0 = ym;   ch[3].y1 = n.y1;
    for (int m : n.members) {
        const int q = ((float)c[m].x < (float)xm ? 0 : 1) + ((float)c[m].y < (float)ym ? 0 : 2);
        ch[q].members.push_back(m);
    }
    for (int q = 0; q < 4; ++q) ch[q].frozen = ch[q].members.size() == 1;
}

// Returns selected candidate indices in the final node-list order.
std::vector<int> distribute_quadtree(const std::vector<Cand>& c, int minX, int maxX, int minY,
                                     int maxY, int N) {
    std::vector<int> result;
    const int nRoots = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nRoots <= 0) return result;
    const float rootW = (float)(maxX - minX) / nRoots;

    std::list<QNode> nodes;
    std::vector<QNode*> roots(nRoots);
    for (int i = 0; i < nRoots; ++i) {
        QNode r;
        r.x0 = (int)(rootW * (float)i);
        r.x1 = (int)(rootW * (float)(i + 1));
        r.y0 = 0;
        r.y1 = maxY - minY;
        nodes.push_back(r);
        roots[i] = &nodes.back();
    }
    for (int i = 0; i < (int)c.size(); ++i) roots[(size_t)((float)c[i].x / rootW)]->members.push_back(i);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->members.size() == 1) { it->frozen = true; ++it; }
        else if (it->members.empty()) it = nodes.erase(it);
        else ++it;
    }

    typedef std::pair<int, QNode*> SizedNode;
    std::vector<SizedNode> expandable;
    auto push_children = [&](QNode ch[4]) {
        int added = 0;
        for (int q = 0; q < 4; ++q) {
            if (ch[q].members.empty()) continue;
            nodes.push_front(ch[q]);
            nodes.front().self = nodes.begin();
            if (ch[q].members.size() > 1) {
                expandable.push_back(SizedNode((int)ch[q].members.size(), &nodes.front()));
                ++added;
            }
        }
        return added;
    };

    bool done = false;
    while (!done) {
        const int before = (int)nodes.size();
        int nExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->frozen) { ++it; continue; }
            QNode ch[4];
            split_node(*it, c, ch);
            nExpand += push_children(ch);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == before) {
            done = true;
        } else if ((int)nodes.size() + 3 * nExpand > N) {
            while (!done) {
                const int before2 = (int)nodes.size();
                std::vector<SizedNode> prev = expandable;
                expandable.clear();
                // canonical: ties in size keep creation order (reference: ties by heap address)
                std::stable_sort(prev.begin(), prev.end(),
                                 [](const SizedNode& a, const SizedNode& b) { return a.first < b.first; });
                for (int j = (int)prev.size() - 1; j >= 0; --j) {
                    QNode ch[4];
                    split_node(*prev[j].second, c, ch);
                    push_children(ch);
                    nodes.erase(prev[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == before2) done = true;
            }
        }
    }

    for (const QNode& n : nodes) {
        int best = n.members[0];
        for (size_t k = 1; k < n.members.size(); ++k)
            if ((float)c[n.members[k]].score > (float)c[best].score) best = n.members[k];
        result.push_back(best);
    }
    return result;
}

// ------------------------------------------------------------------------------------------
struct Level {
    int w = 0, h = 0;
    std::vector<uint8_t> img;      // w*h, tight
    std::vector<uint8_t> blurred;  // w*h
    std::vector<Cand> cand;        // FAST candidates, reference order
    std::vector<int> cand_cell_th; // per candidate: 0 = found at iniTh, 1 = at minTh (diagnostic)
    std::vector<int> selected;     // indices into cand, final list order
    std::vector<float> angle;      // per selected
};

struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    float scaleFactorF;
    double scaleFactor;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> quota;
    std::vector<int> umax;
    int pattern[1024];
    bool have_pattern = false;
    std::vector<Level> lv;
    // outputs
    std::vector<float> kps;  // 6 per kp: x y size angle response octave
    std::vector<uint8_t> desc;

    void init() {
        scaleFactor = (double)scaleFactorF;
        scale.assign(nlevels, 1.f); sigma2.assign(nlevels, 1.f);
        for (int i = 1; i < nlevels; ++i) {
            scale[i] = (float)(scale[i - 1] * scaleFactor);
            sigma2[i] = scale[i] * scale[i];
        }
        invScale.resize(nlevels); invSigma2.resize(nlevels);
        for (int i = 0; i < nlevels; ++i) { invScale[i] = 1.0f / scale[i]; invSigma2[i] = 1.0f / sigma2[i]; }
        quota.resize(nlevels);
        float factor = (float)(1.0f / scaleFactor);
        float per = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; ++l) {
            quota[l] = cvprim::round_half_even(per);
            sum += quota[l];
            per *= factor;
        }
        quota[nlevels - 1] = std::max(nfeatures - sum, 0);
        // circular patch row extents
        umax.assign(kHalfPatch + 1, 0);
        int v, v0;
        const int vmax = cvprim::ifloor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
        const int vmin = cvprim::iceil(kHalfPatch * std::sqrt(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (v = 0; v <= vmax; ++v) umax[v] = cvprim::round_half_even(std::sqrt(hp2 - v * v));
        for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
        lv.resize(nlevels);
    }

    void build_pyramid(const uint8_t* img, int w, int h, size_t stride) {
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            L.w = cvprim::round_half_even((float)w * invScale[l]);
            L.h = cvprim::round_half_even((float)h * invScale[l]);
            L.img.resize((size_t)L.w * L.h);
            if (l == 0) {
                for (int y = 0; y < h; ++y) std::memcpy(&L.img[(size_t)y * w], img + (size_t)y * stride, w);
            } else {
                const Level& P = lv[l - 1];
                cvprim::resize_linear_u8(P.img.data(), P.w, P.h, P.w, L.img.data(), L.w, L.h, L.w);
            }
        }
    }

    void detect_level(int l) {
        Level& L = lv[l];
        L.cand.clear(); L.cand_cell_th.clear(); L.selected.clear(); L.angle.clear();
        const int minBX = kEdge - 3, minBY = minBX;
        const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / kCell), nRows = (int)(height / kCell);
        if (nCols <= 0 || nRows <= 0) return;
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        std::vector<cvprim::FastKp> found;
        for (int i = 0; i < nRows; ++i) {
            const int y0 = minBY + i * hCell;
            int y1 = y0 + hCell + 6;
            if (y0 >= maxBY - 3) continue;
            if (y1 > maxBY) y1 = maxBY;
            for (int j = 0; j < nCols; ++j) {
                const int x0 = minBX + j * wCell;
                int x1 = x0 + wCell + 6;
                if (x0 >= maxBX - 6) continue;
                if (x1 > maxBX) x1 = maxBX;
                const uint8_t* cell = &L.img[(size_t)y0 * L.w + x0];
                int used = 0;
                cvprim::fast9_nms(cell, x1 - x0, y1 - y0, L.w, iniTh, found);
                if (found.empty()) { cvprim::fast9_nms(cell, x1 - x0, y1 - y0, L.w, minTh, found); used = 1; }
                for (const auto& k : found) {
                    L.cand.push_back({k.x + j * wCell, k.y + i * hCell, k.score});
                    L.cand_cell_th.push_back(used);
                }
            }
        }
        L.selected = distribute_quadtree(L.cand, minBX, maxBX, minBY, maxBY, quota[l]);
    }

    float ic_angle(const Level& L, int cx, int cy) const {
        const uint8_t* c = &L.img[(size_t)cy * L.w + cx];
        const int step = L.w;
        int m01 = 0, m10 = 0;
        for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
        for (int v = 1; v <= kHalfPatch; ++v) {
            int vs = 0;
            const int d = umax[v];
            for (int u = -d; u <= d; ++u) {
                const int p = c[u + v * step], m = c[u - v * step];
                vs += p - m;
                m10 += u * (p + m);
            }
            m01 += v * vs;
        }
        return cvprim::fast_atan2_deg((float)m01, (float)m10);
    }

    void describe(const Level& L, int cx, int cy, float angleDeg, uint8_t* out) const {
        const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
        const float ang = angleDeg * factorPI;
        const float a = (float)std::cos(ang), b = (float)std::sin(ang);  // cosf/sinf overloads
        const uint8_t* c = &L.blurred[(size_t)cy * L.w + cx];
        const int step = L.w;
        const int* p = pattern;
        for (int i = 0; i < 32; ++i, p += 32) {
            int val = 0;
            for (int t = 0; t < 8; ++t) {
                const float x0 = (float)p[4 * t], y0 = (float)p[4 * t + 1];
                const float x1 = (float)p[4 * t + 2], y1 = (float)p[4 * t + 3];
                const int r0 = cvprim::round_half_even(x0 * b + y0 * a), c0 = cvprim::round_half_even(x0 * a - y0 * b);
                const int r1 = cvprim::round_half_even(x1 * b + y1 * a), c1 = cvprim::round_half_even(x1 * a - y1 * b);
                val |= (c[r0 * step + c0] < c[r1 * step + c1]) << t;
            }
            out[i] = (uint8_t)val;
        }
    }

    int run(const uint8_t* img, int w, int h, size_t stride) {
        kps.clear(); desc.clear();
        if (!img || w <= 0 || h <= 0) return 0;
        build_pyramid(img, w, h, stride);
        for (int l = 0; l < nlevels; ++l) detect_level(l);
        const int minB = kEdge - 3;
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            L.angle.resize(L.selected.size());
            for (size_t k = 0; k < L.selected.size(); ++k) {
                const Cand& c = L.cand[L.selected[k]];
                L.angle[k] = ic_angle(L, c.x + minB, c.y + minB);
            }
        }
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            if (L.selected.empty()) { L.blurred.clear(); continue; }
            L.blurred.resize((size_t)L.w * L.h);
            cvprim::gaussian7x7_u8(L.img.data(), L.w, L.h, L.w, L.blurred.data(), L.w);
            const int size = (int)(kPatch * scale[l]);
            for (size_t k = 0; k < L.selected.size(); ++k) {
                const Cand& c = L.cand[L.selected[k]];
                const int x = c.x + minB, y = c.y + minB;
                const size_t at = desc.size();
                desc.resize(at + 32);
                if (have_pattern) describe(L, x, y, L.angle[k], &desc[at]);
                float fx = (float)x, fy = (float)y;
                if (l != 0) { fx *= scale[l]; fy *= scale[l]; }
                const float rec[6] = {fx, fy, (float)size, L.angle[k], (float)c.score, (float)l};
                kps.insert(kps.end(), rec, rec + 6);
            }
        }
        return (int)(kps.size() / 6);
    }
};

inline int hamming256(const uint8_t* a, const uint8_t* b) {
    // parallel bit count of the reference, 8 x 32 bit (src/ORBmatcher.cc:1649-1665)
    uint32_t pa[8], pb[8];
    std::memcpy(pa, a, 32); std::memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

}  // namespace

// ==========================================================================================
extern "C" {

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride, uint8_t* dst, int dw,
                          int dh, size_t dstride) {
    cvprim::resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orc_gaussian7x7_u8(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride) {
    cvprim::gaussian7x7_u8(src, w, h, sstride, dst, dstride);
}
void orc_copy_make_border(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst,
                          size_t dstride, int border) {
    cvprim::copy_make_border_reflect101(src, w, h, sstride, dst, dstride, border, border, border, border);
}
// returns count; writes up to cap (x,y,score) int triples
int orc_fast9_nms(const uint8_t* img, int w, int h, size_t stride, int th, int* out, int cap) {
    std::vector<cvprim::FastKp> k;
    cvprim::fast9_nms(img, w, h, stride, th, k);
    for (int i = 0; i < (int)k.size() && i < cap; ++i) { out[3 * i] = k[i].x; out[3 * i + 1] = k[i].y; out[3 * i + 2] = k[i].score; }
    return (int)k.size();
}
int orc_fast_score(const uint8_t* p, size_t stride) { return cvprim::fast_score(p, stride); }
float orc_fast_atan2(float y, float x) { return cvprim::fast_atan2_deg(y, x); }
int orc_round(float v) { return cvprim::round_half_even(v); }

// quadtree on a raw candidate list (x,y relative to min borders; score); returns #selected
int orc_quadtree(const int* xs, const int* ys, const int* scores, int n, int minX, int maxX, int minY,
                 int maxY, int N, int* out_idx, int cap) {
    std::vector<Cand> c(n);
    for (int i = 0; i < n; ++i) c[i] = {xs[i], ys[i], scores[i]};
    std::vector<int> sel = distribute_quadtree(c, minX, maxX, minY, maxY, N);
    for (int i = 0; i < (int)sel.size() && i < cap; ++i) out_idx[i] = sel[i];
    return (int)sel.size();
}

void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                           const int32_t* pattern1024) {
    Extractor* e = new Extractor();
    e->nfeatures = nfeatures; e->scaleFactorF = scaleFactor; e->nlevels = nlevels;
    e->iniTh = iniTh; e->minTh = minTh;
    if (pattern1024) { std::memcpy(e->pattern, pattern1024, sizeof(e->pattern)); e->have_pattern = true; }
    e->init();
    return e;
}
void orc_extractor_destroy(void* h) { delete (Extractor*)h; }
int orc_extract(void* h, const uint8_t* img, int w, int ht, size_t stride) {
    return ((Extractor*)h)->run(img, w, ht, stride);
}
// copy out results of the last orc_extract
void orc_get_result(void* h, float* kps6, uint8_t* desc32) {
    Extractor* e = (Extractor*)h;
    if (kps6) std::memcpy(kps6, e->kps.data(), e->kps.size() * sizeof(float));
    if (desc32) std::memcpy(desc32, e->desc.data(), e->desc.size());
}
void orc_get_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* quota, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; ++i) {
        if (scale) scale[i] = e->scale[i];
        if (invScale) invScale[i] = e->invScale[i];
        if (sigma2) sigma2[i] = e->sigma2[i];
        if (invSigma2) invSigma2[i] = e->invSigma2[i];
        if (quota) quota[i] = e->quota[i];
    }
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
}
void orc_level_dims(void* h, int l, int* w, int* ht, int* ncand, int* nsel) {
    Level& L = ((Extractor*)h)->lv[l];
    *w = L.w; *ht = L.h; *ncand = (int)L.cand.size(); *nsel = (int)L.selected.size();
}
void orc_level_image(void* h, int l, uint8_t* out, int blurred) {
    Level& L = ((Extractor*)h)->lv[l];
    const std::vector<uint8_t>& s = blurred ? L.blurred : L.img;
    std::memcpy(out, s.data(), s.size());
}
int orc_level_has_blur(void* h, int l) { return !((Extractor*)h)->lv[l].blurred.empty(); }
// candidates as (x,y,score) in level coordinates (border added back), reference order
void orc_level_candidates(void* h, int l, int* xys) {
    Level& L = ((Extractor*)h)->lv[l];
    for (size_t i = 0; i < L.cand.size(); ++i) {
        xys[3 * i] = L.cand[i].x + kEdge - 3; xys[3 * i + 1] = L.cand[i].y + kEdge - 3; xys[3 * i + 2] = L.cand[i].score;
    }
}
void orc_level_selected(void* h, int l, int* cand_idx, float* angles) {
    Level& L = ((Extractor*)h)->lv[l];
    for (size_t i = 0; i < L.selected.size(); ++i) {
        if (cand_idx) cand_idx[i] = L.selected[i];
        if (angles) angles[i] = L.angle[i];
    }
}

// ---- matching ---------------------------------------------------------------------------
int orc_hamming(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

// best / second best over all of B for each row of A; strict '<' updates in iteration order
// (the shared selection rule of src/ORBmatcher.cc:104-116, 218-227, 446-457, 588-597).
void orc_knn2(const uint8_t* A, int nA, const uint8_t* B, int nB, int* idx, int* d1, int* d2) {
    for (int i = 0; i < nA; ++i) {
        int b1 = 256, b2 = 256, bi = -1;
        for (int j = 0; j < nB; ++j) {
            const int d = hamming256(A + (size_t)i * 32, B + (size_t)j * 32);
            if (d < b1) { b2 = b1; b1 = d; bi = j; }
            else if (d < b2) b2 = d;
        }
        idx[i] = bi; d1[i] = b1; d2[i] = b2;
    }
}

// same, over per-query candidate lists in CSR form (candidate order = iteration order)
void orc_knn2_lists(const uint8_t* A, int nA, const uint8_t* B, const int* offsets, const int* cands,
                    int* idx, int* d1, int* d2) {
    for (int i = 0; i < nA; ++i) {
        int b1 = 256, b2 = 256, bi = -1;
        for (int k = offsets[i]; k < offsets[i + 1]; ++k) {
            const int j = cands[k];
            const int d = hamming256(A + (size_t)i * 32, B + (size_t)j * 32);
            if (d < b1) { b2 = b1; b1 = d; bi = j; }
            else if (d < b2) b2 = d;
        }
        idx[i] = bi; d1[i] = b1; d2[i] = b2;
    }
}

// ---- stereo ---------------------------------------------------------------------------------
// Frame::ComputeStereoMatches, /root/reference/src/Frame.cc:466-640, on the results + pyramids that the
// two extractor handles hold from their last orc_extract (left, right). uRight / depth: nL floats
// (-1 = no match). Returns the number of matches kept.
int orc_stereo_match(void* hL, void* hR, float mbf, float mb, float* uRight, float* depth) {
    Extractor& EL = *(Extractor*)hL; Extractor& ER = *(Extractor*)hR;
    const int N = (int)(EL.kps.size() / 6), Nr = (int)(ER.kps.size() / 6);
    for (int i = 0; i < N; ++i) { uRight[i] = -1.0f; depth[i] = -1.0f; }
    if (N == 0) return 0;
    const int thOrbDist = (100 + 50) / 2;
    const int nRows = EL.lv[0].h;
    std::vector<std::vector<size_t> > rows(nRows);
    auto KP = [](Extractor& E, int i, int f) { return E.kps[(size_t)i * 6 + f]; };
    for (int iR = 0; iR < Nr; ++iR) {
        const float kpY = KP(ER, iR, 1);
        const float r = 2.0f * ER.scale[(int)KP(ER, iR, 5)];
        const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi) if (yi >= 0 && yi < nRows) rows[yi].push_back(iR);
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int> > distIdx;
    for (int iL = 0; iL < N; ++iL) {
        const int levelL = (int)KP(EL, iL, 5);
        const float vL = KP(EL, iL, 1), uL = KP(EL, iL, 0);
        const std::vector<size_t>& cand = rows[(size_t)vL];
        if (cand.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100;
        size_t bestIdxR = 0;
        for (size_t c = 0; c < cand.size(); ++c) {
            const size_t iR = cand[c];
            const int oR = (int)KP(ER, (int)iR, 5);
            if (oR < levelL - 1 || oR > levelL + 1) continue;
            const float uR = KP(ER, (int)iR, 0);
            if (uR >= minU && uR <= maxU) {
                const int d = hamming256(&EL.desc[(size_t)iL * 32], &ER.desc[iR * 32]);
                if (d < bestDist) { bestDist = d; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = KP(ER, (int)bestIdxR, 0);
            const float sf = EL.invScale[levelL];
            const float scaleduL = std::round(uL * sf), scaledvL = std::round(vL * sf), scaleduR0 = std::round(uR0 * sf);
            const int w = 5, L = 5;
            const Level& PL = EL.lv[levelL]; const Level& PR = ER.lv[levelL];
            const int cuL = (int)scaleduL, cvL = (int)scaledvL, cuR = (int)scaleduR0;
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= PR.w) continue;
            int best = 2147483647, bestinc = 0;
            float vD[2 * 5 + 1];
            // pixels outside the level come from the reference's 19 px BORDER_REFLECT_101 frame (1122-1128)
            auto px = [](const Level& P, int x, int y) { return (int)P.img[(size_t)cvprim::reflect101(y, P.h) * P.w + cvprim::reflect101(x, P.w)]; };
            const int cL = px(PL, cuL, cvL);
            for (int inc = -L; inc <= L; ++inc) {
                const int cR = px(PR, cuR + inc, cvL);
                float dist = 0;
                for (int dy = -w; dy <= w; ++dy)
                    for (int dx = -w; dx <= w; ++dx) {
                        const float a = (float)px(PL, cuL + dx, cvL + dy) - (float)cL;
                        const float b = (float)px(PR, cuR + inc + dx, cvL + dy) - (float)cR;
                        dist += std::fabs(a - b);
                    }
                if (dist < best) { best = (int)dist; bestinc = inc; }
                vD[L + inc] = dist;
            }
            if (bestinc == -L || bestinc == L) continue;
            const float d1 = vD[L + bestinc - 1], d2 = vD[L + bestinc], d3 = vD[L + bestinc + 1];
            const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = EL.scale[levelL] * ((float)scaleduR0 + (float)bestinc + deltaR);
            float disparity = uL - bestuR;
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                distIdx.push_back(std::pair<int, int>(best, iL));
            }
        }
    }
    if (distIdx.empty()) return 0;  // (the reference indexes an empty vector here)
    std::sort(distIdx.begin(), distIdx.end());
    const float median = (float)distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)distIdx.size();
    for (int i = (int)distIdx.size() - 1; i >= 0; --i) {
        if ((float)distIdx[i].first < thDist) break;
        uRight[distIdx[i].second] = -1; depth[distIdx[i].second] = -1; --kept;
    }
    return kept;
}
}  // extern "C"
