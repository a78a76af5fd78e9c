"""Validates the oracle restatement (oracle/orb_oracle.cc) against the reference's OWN
ORBextractor.cc compiled unmodified (oracle/_ref, built by oracle/build_ref.sh against the cv2-
pinned OpenCV stand-in). Bit-exact keypoints + descriptors for the canonical tie-break build; the
verbatim (heap-address tie-break) build is compared as a set and the differing fraction bounded."""
import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import synth

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (reference tree absent)")

CONFIGS = [  # w, h, nfeatures, iniTh  (TUM mono / KITTI / EuRoC settings)
    (640, 480, 1000, 20),
    (1241, 376, 2000, 20),
    (752, 480, 1200, 20),
    (1241, 376, 2000, 12),
]


@pytest.mark.parametrize("w,h,nf,ini", CONFIGS)
@pytest.mark.parametrize("kind,seed", [("blocks", 0), ("blocks", 1), ("blurnoise", 0), ("noise", 2)])
def test_oracle_equals_reference_canonical(w, h, nf, ini, kind, seed):
    img = synth.image(kind, w, h, seed)
    k, d = O.OracleExtractor(nf, 1.2, 8, ini, 7)(img)
    ref = O.RefExtractor(nf, 1.2, 8, ini, 7, kind="canonical")
    k2, d2 = ref(img)
    assert len(k) == len(k2) and len(k) > nf // 2
    assert np.array_equal(k.view(np.uint32), k2.view(np.uint32))  # bit patterns, incl. angles
    assert np.array_equal(d, d2)


def test_tables_and_pyramid_equal_reference():
    img = synth.image("blurnoise", 752, 480, 3)
    o = O.OracleExtractor(1200, 1.2, 8)
    o(img)
    r = O.RefExtractor(1200, 1.2, 8)
    r(img)
    to, tr = o.tables(), r.tables()
    for key in tr:
        assert np.array_equal(to[key].view(np.uint32), tr[key].view(np.uint32)), key
    assert to["quota"].tolist() == [261, 217, 181, 151, 126, 105, 87, 72]
    assert to["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    for l in range(8):
        assert np.array_equal(o.level(l)["img"], r.level_image(l))


def test_other_pyramid_settings():
    img = synth.image("blocks", 640, 480, 4)
    for nf, s, nl, ini, mn in [(500, 1.2, 8, 20, 7), (1500, 1.1, 5, 15, 5), (300, 1.5, 4, 30, 10), (2000, 1.2, 8, 20, 7)]:
        k, d = O.OracleExtractor(nf, s, nl, ini, mn)(img)
        k2, d2 = O.RefExtractor(nf, s, nl, ini, mn)(img)
        assert np.array_equal(k.view(np.uint32), k2.view(np.uint32)) and np.array_equal(d, d2)


def test_flat_image_gives_no_keypoints():
    img = synth.image("flat", 640, 480, 0)
    k, d = O.OracleExtractor()(img)
    k2, d2 = O.RefExtractor()(img)
    assert len(k) == 0 and len(k2) == 0


@pytest.mark.skipif(not O.ref_available("verbatim"), reason="verbatim build absent")
def test_verbatim_tiebreak_differs_only_slightly():
    """The unmodified reference orders equal-size nodes by heap address; the canonical rule may
    pick a different subset. SURVEY.md measured ~1.3 %; bound it at 5 % here and report."""
    tot = diff = 0
    for seed in range(3):
        img = synth.image("blocks", 640, 480, seed)
        k, _ = O.OracleExtractor()(img)
        k2, _ = O.RefExtractor(kind="verbatim")(img)
        a, b = set(map(tuple, k[:, :2].tolist())), set(map(tuple, k2[:, :2].tolist()))
        tot += len(a)
        diff += len(a - b)
    print("canonical-vs-verbatim differing keypoints: %d / %d = %.2f%%" % (diff, tot, 100.0 * diff / tot))
    assert diff / tot < 0.05


@pytest.mark.skipif(not O.matcher_bits_available(), reason="libref_matcher_bits.so absent")
def test_descriptor_distance_and_three_maxima_equal_reference():
    """src/ORBmatcher.cc:1603-1665 compiled from the reference file, against the oracle restatement and the
    product's host mirrors (ORBmatcher.DescriptorDistance / ComputeThreeMaxima)."""
    from multiagent_orb_slam2_b200.matcher import ORBmatcher
    ref = O.RefMatcherBits()
    rng = np.random.default_rng(5)
    D = rng.integers(0, 256, (400, 32), dtype=np.uint8)
    D[0] = 0; D[1] = 255; D[2] = D[3]
    for i in range(0, 400, 2):
        want = ref.descriptor_distance(D[i], D[i + 1])
        assert O.hamming(D[i], D[i + 1]) == want
        assert ORBmatcher.DescriptorDistance(D[i], D[i + 1]) == want
    assert ref.descriptor_distance(D[0], D[1]) == 256
    cases = [rng.integers(0, hi, 30) for hi in (1, 2, 5, 50, 1000) for _ in range(40)]
    cases += [np.zeros(30, int), np.full(30, 7), np.arange(30), np.arange(30)[::-1], [100, 10, 9] + [0] * 27,
              [100, 9, 9] + [0] * 27, [100, 10, 10] + [0] * 27, [10, 1, 0] + [0] * 27, [11, 1, 1] + [0] * 27]
    for s in cases:
        want = ref.three_maxima(s)
        assert O.three_maxima([int(x) for x in s]) == want, list(s)
        assert ORBmatcher.ComputeThreeMaxima(s) == want, list(s)
