// TEST INFRASTRUCTURE ONLY (oracle side). Array-form restatement of DistributeOctTree
// (/root/reference/src/ORBextractor.cc:539-763) used to validate, on the CPU, the formulation the
// CUDA kernel implements (multiagent_orb_slam2_b200/csrc/quadtree.cuh):
//
//   * every candidate gets a path key: root index, then 2 bits per depth (bit0 = right half,
//     bit1 = bottom half) from the recursive ceil-halving of DivideNode (481-537);
//   * candidates are sorted by key, so every tree node is a contiguous range [lo,hi) + a depth;
//   * the std::list is an array in list order; one "step" processes a sequence P of expandable
//     nodes (list order for a normal pass; (size desc, list position asc) for the final phase,
//     which is the canonical tie-break), pushes each one's children to the front (n1..n4, so the
//     front reads n4..n1) and may stop early once the list holds >= N nodes (730-731).
//
// Checked against distribute_quadtree() of orb_oracle.cc by tests/test_quadtree_arrayform.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace {

struct Node { int lo, hi, depth; };

constexpr int kMaxDepth = 13;

inline uint32_t path_key(int x, int y, int minX, int maxX, int minY, int maxY, int nRoots, float rootW) {
    const int r = (int)((float)x / rootW);
    int x0 = (int)(rootW * (float)r), x1 = (int)(rootW * (float)(r + 1));
    int y0 = 0, y1 = maxY - minY;
    (void)minX; (void)maxX; (void)nRoots;
    uint32_t key = (uint32_t)r;
    for (int d = 0; d < kMaxDepth; ++d) {
        const int xm = x0 + ((x1 - x0 + 1) >> 1), ym = y0 + ((y1 - y0 + 1) >> 1);
        const uint32_t xb = x >= xm, yb = y >= ym;
        if (xb) x0 = xm; else x1 = xm;
        if (yb) y0 = ym; else y1 = ym;
        key = (key << 2) | (yb << 1) | xb;
    }
    return key;
}

inline uint32_t digit_at(uint32_t key, int depth) {  // depth 1..kMaxDepth
    return (key >> (2 * (kMaxDepth - depth))) & 3u;
}

}  // namespace

extern "C" int orc_quadtree_arrayform(const int* xs, const int* ys, const int* scores, int n, int minX, int maxX,
                                      int minY, int maxY, int N, int* out_idx, int cap) {
    const int nRoots = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nRoots <= 0 || n == 0) return 0;
    const float rootW = (float)(maxX - minX) / nRoots;

    std::vector<uint32_t> key(n);
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) { key[i] = path_key(xs[i], ys[i], minX, maxX, minY, maxY, nRoots, rootW); order[i] = i; }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] != key[b] ? key[a] < key[b] : a < b; });
    std::vector<uint32_t> sk(n);
    for (int i = 0; i < n; ++i) sk[i] = key[order[i]];

    // roots
    std::vector<Node> list;
    for (int r = 0; r < nRoots; ++r) {
        const uint32_t lo_key = (uint32_t)r << (2 * kMaxDepth), hi_key = (uint32_t)(r + 1) << (2 * kMaxDepth);
        const int lo = (int)(std::lower_bound(sk.begin(), sk.end(), lo_key) - sk.begin());
        const int hi = (int)(std::lower_bound(sk.begin(), sk.end(), hi_key) - sk.begin());
        if (hi > lo) list.push_back({lo, hi, 0});
    }

    bool finalPhase = false, done = false;
    while (!done) {
        const int before = (int)list.size();
        // expandable nodes, in processing order
        std::vector<int> P;
        for (int i = 0; i < before; ++i)
            if (list[i].hi - list[i].lo >= 2 && list[i].depth < kMaxDepth) P.push_back(i);
        if (finalPhase)
            std::stable_sort(P.begin(), P.end(), [&](int a, int b) {
                return (list[a].hi - list[a].lo) > (list[b].hi - list[b].lo);
            });  // ties: list position ascending (stable over ascending P)
        // children of every expandable node
        const int m = (int)P.size();
        std::vector<int> cb(5 * (size_t)m);  // child boundaries
        std::vector<int> cc(m);              // non-empty children
        for (int k = 0; k < m; ++k) {
            const Node& nd = list[P[k]];
            int* b = &cb[5 * (size_t)k];
            b[0] = nd.lo; b[4] = nd.hi;
            for (int q = 1; q < 4; ++q) {
                int lo = nd.lo, hi = nd.hi;  // first index with digit >= q
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)digit_at(sk[mid], nd.depth + 1) >= q) hi = mid; else lo = mid + 1; }
                b[q] = lo;
            }
            cc[k] = 0;
            for (int q = 0; q < 4; ++q) cc[k] += b[q + 1] > b[q];
        }
        // how many get processed
        int processed = m, size = before;
        if (finalPhase) {
            processed = 0;
            for (int k = 0; k < m; ++k) { size += cc[k] - 1; processed = k + 1; if (size >= N) break; }
        } else {
            for (int k = 0; k < m; ++k) size += cc[k] - 1;
        }
        // new list: children of P[processed-1] (q=3..0) ... children of P[0], then the untouched old nodes
        std::vector<char> gone(before, 0);
        std::vector<Node> next;
        int nExpand = 0;
        for (int k = processed - 1; k >= 0; --k) {
            const Node& nd = list[P[k]];
            gone[P[k]] = 1;
            const int* b = &cb[5 * (size_t)k];
            for (int q = 3; q >= 0; --q)
                if (b[q + 1] > b[q]) { next.push_back({b[q], b[q + 1], nd.depth + 1}); nExpand += (b[q + 1] - b[q]) >= 2; }
        }
        for (int i = 0; i < before; ++i) if (!gone[i]) next.push_back(list[i]);
        list.swap(next);
        const int after = (int)list.size();
        if (after >= N || after == before) done = true;
        else if (!finalPhase && after + 3 * nExpand > N) finalPhase = true;
    }

    int cnt = 0;
    for (const Node& nd : list) {
        int best = order[nd.lo];
        for (int i = nd.lo + 1; i < nd.hi; ++i) {
            const int c = order[i];
            if (scores[c] > scores[best] || (scores[c] == scores[best] && c < best)) best = c;
        }
        if (cnt < cap) out_idx[cnt] = best;
        ++cnt;
    }
    return cnt;
}
