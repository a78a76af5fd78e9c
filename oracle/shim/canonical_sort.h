// TEST INFRASTRUCTURE ONLY (oracle). Force-included (-include) when building the CANONICAL
// variant of the reference extractor: DistributeOctTree's unqualified
//   sort(vPrevSizeAndPointerToNode.begin(), vPrevSizeAndPointerToNode.end())
// (/root/reference/src/ORBextractor.cc:684) orders pair<int,ExtractorNode*> and so breaks ties in
// node size by heap address. This non-template overload is found by argument-dependent lookup
// (ExtractorNode lives in ORB_SLAM2) and wins over std::sort; it orders by size only and keeps
// creation order among ties - the documented canonical tie-break. No reference line changes.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>
namespace ORB_SLAM2 {
class ExtractorNode;
typedef std::vector<std::pair<int, ExtractorNode*> >::iterator SizedNodeIt;
inline void sort(SizedNodeIt first, SizedNodeIt last) {
    std::stable_sort(first, last, [](const std::pair<int, ExtractorNode*>& a,
                                     const std::pair<int, ExtractorNode*>& b) { return a.first < b.first; });
}
}  // namespace ORB_SLAM2
