#!/bin/bash
# TEST INFRASTRUCTURE ONLY (oracle). Builds oracle/_ref/ from the reference sources WHERE THEY
# LIE under /root/reference (nothing is copied into the repo; outputs are git-ignored):
#   libref_orb_verbatim.so   src/ORBextractor.cc unmodified (node-size ties by heap address)
#   libref_orb_canonical.so  same file, plus the ADL sort overload of shim/canonical_sort.h
# Both link the cv2-pinned OpenCV stand-in of oracle/shim. Flags follow the reference
# (CMakeLists.txt:10-11: -O3 -march=native) except that FMA contraction is pinned off so that the
# float rotation in computeOrbDescriptor is compiler-independent.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${ORB_REFERENCE_ROOT:-/root/reference}"
if [ ! -f "$ref/src/ORBextractor.cc" ]; then
  echo "build_ref: $ref not present, keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$here/_ref"
common=(-std=c++14 -O3 -march=x86-64-v3 -ffp-contract=off -fPIC -shared -w
        -I"$here/shim" -I"$ref/include" "$ref/src/ORBextractor.cc" "$here/ref_wrap.cc")
g++ "${common[@]}" -o "$here/_ref/libref_orb_verbatim.so"
g++ -include "$here/shim/canonical_sort.h" "${common[@]}" -o "$here/_ref/libref_orb_canonical.so"
#   libref_dbow.so           the vendored DBoW2 (FORB, BowVector, FeatureVector, ScoringObject,
#                            TemplatedVocabulary.h) unmodified + ref_dbow_wrap.cc: ORBVocabulary::transform
dbow="$ref/Thirdparty/DBoW2"
g++ -std=c++14 -O3 -march=x86-64-v3 -ffp-contract=off -fPIC -shared -w -I"$here/shim" -I"$dbow" \
    "$dbow/DBoW2/FORB.cpp" "$dbow/DBoW2/BowVector.cpp" "$dbow/DBoW2/FeatureVector.cpp" "$dbow/DBoW2/ScoringObject.cpp" \
    "$dbow/DUtils/Random.cpp" "$dbow/DUtils/Timestamp.cpp" "$here/ref_dbow_wrap.cc" -o "$here/_ref/libref_dbow.so"
#   libref_matcher_bits.so   the two self-contained members of src/ORBmatcher.cc - ComputeThreeMaxima (1603-1644) and
#                            DescriptorDistance (1647-1665) - piped straight from the reference file into the compiler
#                            between a stub class declaration and C wrappers (the rest of ORBmatcher.cc needs the SLAM graph
#                            types and cannot be built here). Nothing of the reference is written to disk.
{
  cat <<'PRE'
#include <cstdint>
#include <vector>
#include "cvshim.h"
using namespace std;
namespace ORB_SLAM2 {
class ORBmatcher {
public:
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3);
};
PRE
  sed -n '1603,1665p' "$ref/src/ORBmatcher.cc"
  cat <<'POST'
}  // namespace ORB_SLAM2
extern "C" int refm_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma(1, 32, CV_8U, (void*)a, 32), mb(1, 32, CV_8U, (void*)b, 32);
    return ORB_SLAM2::ORBmatcher::DescriptorDistance(ma, mb);
}
extern "C" void refm_three_maxima(const int* sizes, int L, int* out3) {
    std::vector<std::vector<int> > h(L);
    for (int i = 0; i < L; ++i) h[i].resize(sizes[i]);
    int a = -1, b = -1, c = -1;
    ORB_SLAM2::ORBmatcher m;
    m.ComputeThreeMaxima(h.data(), L, a, b, c);
    out3[0] = a; out3[1] = b; out3[2] = c;
}
POST
} | g++ -std=c++14 -O3 -march=x86-64-v3 -fPIC -shared -w -I"$here/shim" -x c++ - -o "$here/_ref/libref_matcher_bits.so"
#   libref_slam.so           the reference's matcher and map data model, UNMODIFIED: src/ORBmatcher.cc, Frame.cc, KeyFrame.cc,
#                            MapPoint.cc, Map.cc, KeyFrameDatabase.cc, ORBextractor.cc + the vendored DBoW2, against the
#                            stand-in (float cv::Mat expressions pinned to cv2) + ref_slam_wrap.cc. The extractor inside uses
#                            the canonical node-size tie-break like libref_orb_canonical.so. Eigen / g2o / Pangolin are not
#                            needed: Converter.h is the only header that pulls them (see shim/slam_preamble.h).
g++ -std=c++14 -O2 -march=x86-64-v3 -ffp-contract=off -fPIC -shared -w -pthread \
    -include "$here/shim/canonical_sort.h" -include "$here/shim/slam_preamble.h" -I"$here/shim" -I"$ref/include" -I"$ref" \
    "$ref/src/ORBmatcher.cc" "$ref/src/Frame.cc" "$ref/src/KeyFrame.cc" "$ref/src/MapPoint.cc" "$ref/src/Map.cc" \
    "$ref/src/KeyFrameDatabase.cc" "$ref/src/ORBextractor.cc" \
    "$dbow/DBoW2/FORB.cpp" "$dbow/DBoW2/BowVector.cpp" "$dbow/DBoW2/FeatureVector.cpp" "$dbow/DBoW2/ScoringObject.cpp" \
    "$dbow/DUtils/Random.cpp" "$dbow/DUtils/Timestamp.cpp" "$here/ref_slam_wrap.cc" -o "$here/_ref/libref_slam.so"
#   real_types_test         tests/cpp/real_types_test.cc: the PRODUCT's C++ facade templates (include/orbslam2_b200/ORBmatcher.h)
#                            instantiated on the reference's own Frame / KeyFrame / MapPoint classes, beside the reference's
#                            ORBmatcher on identical objects. Needs the reference headers, so it is built here and travels
#                            to the GPU box as a binary; run by tests/test_gpu_real_types.py.
repo="$(cd "$here/.." && pwd)"
if [ -f "$repo/multiagent_orb_slam2_b200/lib/liborb_b200.so" ]; then
  g++ -std=c++14 -O2 -march=x86-64-v3 -ffp-contract=off -w -pthread \
      -include "$here/shim/slam_preamble.h" -I"$here/shim" -I"$ref/include" -I"$ref" -I"$repo/include" \
      "$repo/tests/cpp/real_types_test.cc" -o "$here/_ref/real_types_test" \
      -L"$here/_ref" -lref_slam -L"$repo/multiagent_orb_slam2_b200/lib" -lorb_b200 \
      -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../multiagent_orb_slam2_b200/lib'
else
  echo "build_ref: liborb_b200.so not built yet, skipping real_types_test" >&2
fi
echo "build_ref: built $(ls "$here/_ref" | tr '\n' ' ')"
