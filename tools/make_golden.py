"""Generates tests/golden/*.npz: outputs of the reference's own ORBextractor.cc (oracle/_ref, canonical
tie-break build, compiled here from /root/reference against the cv2-pinned OpenCV stand-in) on seeded
synthetic images, plus Hamming kNN-2 results of the oracle loop on seeded descriptor sets. The fixtures
travel to the GPU box (where /root/reference does not exist). Run from the repo root:
    python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from multiagent_orb_slam2_b200 import synth  # noqa: E402

CASES = [  # name, kind, w, h, seed, nfeatures, scale, nlevels, ini, min
    ("tum_blocks", "blocks", 640, 480, 0, 1000, 1.2, 8, 20, 7),
    ("tum_blurnoise", "blurnoise", 640, 480, 1, 1000, 1.2, 8, 20, 7),
    ("kitti_blocks", "blocks", 1241, 376, 2, 2000, 1.2, 8, 20, 7),
    ("euroc_noise", "noise", 752, 480, 3, 1200, 1.2, 8, 20, 7),
    ("small_odd", "blocks", 321, 243, 4, 300, 1.3, 4, 15, 5),
]

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
assert O.ref_available("canonical"), "build oracle/_ref first (oracle/build_ref.sh)"
for name, kind, w, h, seed, nf, sc, nl, ini, mn in CASES:
    img = synth.image(kind, w, h, seed)
    k, d = O.RefExtractor(nf, sc, nl, ini, mn, kind="canonical")(img)
    ko, do = O.OracleExtractor(nf, sc, nl, ini, mn)(img)
    assert np.array_equal(k.view(np.uint32), ko.view(np.uint32)) and np.array_equal(d, do), name
    np.savez_compressed(os.path.join(out, "extract_%s.npz" % name), kind=kind, w=w, h=h, seed=seed, nfeatures=nf, scale=sc,
                        nlevels=nl, ini=ini, mn=mn, image_sha1=np.frombuffer(__import__("hashlib").sha1(img.tobytes()).digest(), np.uint8),
                        keypoints=k, descriptors=d)
    print(name, len(k), "keypoints")

B = synth.descriptors(700, 21)
A = synth.descriptors(500, 22, dup_from=B)
B[350] = B[3]
idx, d1, d2 = O.knn2(A, B)
np.savez_compressed(os.path.join(out, "knn2_500x700.npz"), A=A, B=B, idx=idx, best=d1, second=d2)
print("knn2 fixture written")

# The nine guided searches + Frame::isInFrustum of the reference's OWN ORBmatcher.cc / Frame.cc / KeyFrame.cc / MapPoint.cc
# (oracle/_ref/libref_slam.so, compiled unmodified) on the scenario of tests/guided_scenario.py: the GPU box compares the
# C++ facade with these.
import guided_reference as GR  # noqa: E402
import guided_scenario as G  # noqa: E402
import ref_slam  # noqa: E402

assert ref_slam.available(), "build oracle/_ref first (oracle/build_ref.sh)"
for seed in (0, 1):
    views, shift, scale = G.extract_pair_cpu(seed)
    sc = G.make_scenario(seed, views, shift, scale)
    res, fuse, tail, frustum = GR.reference_results(sc)
    arrays = {"r%02d" % i: a for i, a in enumerate(res)}
    arrays.update({"t%02d" % i: a for i, a in enumerate(tail)})
    arrays.update({"fuse_" + k: np.asarray(v) for k, v in fuse.items()})
    np.savez_compressed(os.path.join(out, "guided_reference_seed%d.npz" % seed), frustum=frustum, k0=views[0][0], k1=views[1][0],
                        shift=np.array(shift), **arrays)
    print("guided reference fixture", seed, [int(a[0]) for a in res[::2]])
