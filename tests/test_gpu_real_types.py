"""Rows b-2 / b-3 of SURVEY.md section 8: the product's C++ facade (include/orbslam2_b200/ORBmatcher.h, eleven search
templates) instantiated on the REFERENCE'S OWN Frame / KeyFrame / MapPoint classes, next to the reference's own ORBmatcher
(src/ORBmatcher.cc) on identically built objects. oracle/_ref/real_types_test (tests/cpp/real_types_test.cc, built here by
oracle/build_ref.sh because it needs the reference headers; the binary travels to the GPU box) runs every search twice and
dumps both result sets; they must be identical - match sets, and for Fuse the map graph state MapPoint::Replace /
AddObservation leave behind."""
import os
import subprocess

import numpy as np
import pytest

import guided_scenario as G
from test_gpu_cpp_guided import read_results, write_scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "real_types_test")
pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/real_types_test not built (reference tree absent)")

NAMES = ["a10 n", "a10 map", "a11 stereo n", "a11 stereo map", "a11 mono n", "a11 mono map", "a12 cur-kf n", "a12 cur-kf map", "a12 kf-scw n",
         "a12 kf-scw map", "bow kf-f n", "bow kf-f map", "tri n", "tri pairs", "tri-stereo n", "tri-stereo pairs", "fuse n", "fuse held", "fuse bad",
         "fuse nobs", "fuse replaced", "fuse-scw n", "fuse-scw replace", "fuse-scw map", "sim3 n", "sim3 map", "bow kf-kf n", "bow kf-kf map",
         "init n", "init map"]


def run(tmp, seed, facade):
    views, shift, scale = G.extract_pair_cpu(seed)
    sc = G.make_scenario(seed, views, shift, scale)
    spath = os.path.join(tmp, "scenario.bin")
    write_scenario(spath, sc)
    out_ref, out_b200 = os.path.join(tmp, "ref.bin"), os.path.join(tmp, "b200.bin")
    subprocess.check_call([EXE, spath, out_ref, out_b200 if facade else "-"])
    return sc, read_results(out_ref), read_results(out_b200) if facade else None


def test_reference_half_of_the_driver_equals_oracle(tmp_path):
    """CPU: the driver's reference pass (no device needed) reproduces the oracle, so the scenario file and the object
    construction in the driver are the ones tests/test_oracle_vs_reference_matcher.py pins."""
    sc, ref, _ = run(str(tmp_path), 0, False)
    want = G.oracle_results(sc)
    assert len(ref) == len(NAMES)
    for i in list(range(16)):
        assert np.array_equal(ref[i], want[i]), NAMES[i]
    for i, j in zip(range(21, 26), range(19, 24)):
        assert np.array_equal(ref[i], want[j]), NAMES[i]
    assert ref[26][0] >= 10 and ref[28][0] >= 20


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_facade_on_reference_types_equals_reference_matcher(tmp_path, seed):
    sc, ref, got = run(str(tmp_path), seed, True)
    assert len(ref) == len(got) == len(NAMES)
    for name, r, g in zip(NAMES, ref, got):
        assert np.array_equal(r, g), (name, r[:20], g[:20])
    for key, least in [("a10 n", 100), ("a11 stereo n", 50), ("a12 cur-kf n", 50), ("a12 kf-scw n", 30), ("bow kf-f n", 50), ("tri n", 10),
                       ("fuse n", 30), ("fuse-scw n", 30), ("sim3 n", 10), ("bow kf-kf n", 10), ("init n", 20)]:
        assert ref[NAMES.index(key)][0] >= least, (key, ref[NAMES.index(key)][0])
