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
echo "build_ref: built $(ls "$here/_ref")"
