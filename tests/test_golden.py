"""Golden fixtures (tests/golden, written by tools/make_golden.py from the reference's own extractor
compiled here): the oracle must reproduce them on CPU, the CUDA path on the GPU box."""
import glob
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXTRACT = sorted(glob.glob(os.path.join(GOLD, "extract_*.npz")))


def load(path):
    z = np.load(path)
    img = synth.image(str(z["kind"]), int(z["w"]), int(z["h"]), int(z["seed"]))
    assert hashlib.sha1(img.tobytes()).digest() == z["image_sha1"].tobytes(), "synthetic image generator drifted"
    cfg = (int(z["nfeatures"]), float(z["scale"]), int(z["nlevels"]), int(z["ini"]), int(z["mn"]))
    return img, cfg, z["keypoints"], z["descriptors"]


def test_fixtures_present():
    assert len(EXTRACT) >= 5


@pytest.mark.parametrize("path", EXTRACT, ids=[os.path.basename(p) for p in EXTRACT])
def test_oracle_reproduces_golden(path):
    img, cfg, k, d = load(path)
    ko, do = O.OracleExtractor(*cfg)(img)
    assert np.array_equal(ko.view(np.uint32), k.view(np.uint32)) and np.array_equal(do, d)


def test_oracle_knn2_reproduces_golden():
    z = np.load(os.path.join(GOLD, "knn2_500x700.npz"))
    idx, d1, d2 = O.knn2(z["A"], z["B"])
    assert np.array_equal(idx, z["idx"]) and np.array_equal(d1, z["best"]) and np.array_equal(d2, z["second"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", EXTRACT, ids=[os.path.basename(p) for p in EXTRACT])
def test_cuda_reproduces_golden(path):
    from multiagent_orb_slam2_b200 import ORBextractor
    img, cfg, k, d = load(path)
    gk, gd = ORBextractor(*cfg)(img)
    got = np.stack([gk["x"], gk["y"], gk["size"], gk["angle"], gk["response"], gk["octave"].astype(np.float32)], 1)
    assert np.array_equal(got.view(np.uint32), k.view(np.uint32)) and np.array_equal(gd, d)


@pytest.mark.gpu
def test_cuda_knn2_reproduces_golden():
    from multiagent_orb_slam2_b200 import ORBmatcher
    z = np.load(os.path.join(GOLD, "knn2_500x700.npz"))
    idx, d1, d2 = ORBmatcher().knn2(z["A"], z["B"])
    assert np.array_equal(idx, z["idx"]) and np.array_equal(d1, z["best"]) and np.array_equal(d2, z["second"])


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_reproduces_guided_reference_golden(seed):
    """tests/golden/guided_reference_seed*.npz: the nine guided searches + Frame::isInFrustum as computed by the reference's
    own ORBmatcher.cc / Frame.cc (oracle/_ref/libref_slam.so) - the oracle restatements must reproduce them."""
    import guided_scenario as G
    z = np.load(os.path.join(GOLD, "guided_reference_seed%d.npz" % seed))
    views, shift, scale = G.extract_pair_cpu(seed)
    assert np.array_equal(views[0][0].view(np.uint32), z["k0"].view(np.uint32)) and tuple(z["shift"]) == tuple(shift)
    sc = G.make_scenario(seed, views, shift, scale)
    want = G.oracle_results(sc)
    for i in range(16):
        assert np.array_equal(want[i], z["r%02d" % i]), G.NAMES[i]
    for i in range(5):
        assert np.array_equal(want[19 + i], z["t%02d" % i]), G.NAMES[19 + i]
    fr, alive = sc["frustum"], sc["frustum"]["alive"].astype(bool)
    assert np.array_equal(z["frustum"][:, 0] != 0, alive)
    for col, key in ((1, "u"), (2, "v"), (3, "ur"), (5, "view_cos")):
        assert np.array_equal(z["frustum"][alive, col].view(np.uint32), fr[key][alive].view(np.uint32)), key
