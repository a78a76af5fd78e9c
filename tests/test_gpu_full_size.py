"""Parity at BASELINE.json's full sizes, where the oracle cannot run everything in seconds: size-independent
properties plus full oracle checks on random samples."""
import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import ORBextractor, ORBmatcher, synth

pytestmark = pytest.mark.gpu


def test_cross_map_200k_sets_sampled_against_oracle():
    """BASELINE config 5: 200 000 x 200 000 descriptors. Full oracle check on 96 random query rows, self-match
    identity on the whole set, and agreement of every reported best distance with a recomputation."""
    n = 200_000
    base = synth.descriptors(n, 11)
    A = synth.descriptors_fast(n, 12, base, 60)
    B = synth.descriptors_fast(n, 13, base, 60)
    m = ORBmatcher(0.75)
    idx, d1, d2 = m.knn2(A, B)
    rng = np.random.default_rng(0)
    rows = rng.choice(n, 96, replace=False)
    oi, o1, o2 = O.knn2(A[rows], B, threads=8)
    assert np.array_equal(idx[rows], oi) and np.array_equal(d1[rows], o1) and np.array_equal(d2[rows], o2)
    # every reported best distance is the true distance to the reported row, and second >= best
    recomputed = np.unpackbits(A ^ B[idx], axis=1).sum(1)
    assert np.array_equal(recomputed, d1) and (d2 >= d1).all() and (idx >= 0).all()
    # matching a set against itself: every row finds itself at distance 0 (first minimum = lowest index of a duplicate)
    si, s1, _ = m.knn2(B, B)
    assert (s1 == 0).all() and (si <= np.arange(n)).all()
    dup = np.flatnonzero(si != np.arange(n))
    assert all((B[i] == B[si[i]]).all() for i in dup[:50])
    acc = m.accept(idx, d1, d2)
    assert 0.2 < (acc >= 0).mean() < 0.95


def test_euroc_batch_of_64_frames():
    """BASELINE config 3: 752 x 480, 1200 keypoints, 64 frames per launch, full frame-to-frame matching. Every
    8th frame is checked completely against the oracle, all frames through invariants."""
    from multiagent_orb_slam2_b200.frontend import AgentFrontend
    frames = []
    for s in range(8):
        a, b, _ = synth.shifted_pair("blocks", 752, 480, 40 + s)
        frames += [a, b]
    imgs = np.stack(frames * 4)
    fe = AgentFrontend(752, 480, nfeatures=1200, max_batch=64)
    kps, desc, counts, match = fe.process(imgs)
    assert counts.min() > 1100 and counts.max() <= fe.cap
    o = O.OracleExtractor(1200, 1.2, 8, 20, 7)
    for i in range(0, 64, 8):
        ok, od = o(imgs[i])
        assert counts[i] == len(ok) and np.array_equal(desc[i, :counts[i]], od)
        ok2, od2 = o(imgs[(i + 1) % 64])
        oi, o1, o2 = O.knn2(od, od2)
        ref = ORBmatcher(0.9).accept(oi, o1, o2, th=50, inclusive=True)
        assert np.array_equal(match[i, :counts[i]], ref)
    # identical input frames give identical outputs anywhere in the batch (frame 0 == frame 16 == frame 32 ...)
    for i in range(16, 64, 16):
        assert counts[i] == counts[0] and np.array_equal(desc[i, :counts[i]], desc[0, :counts[0]])
        assert np.array_equal(kps[i, :counts[i]], kps[0, :counts[0]])
    # octaves ascend, keypoints stay inside the image, at most nfeatures + 3 per level overshoot
    for i in range(64):
        k = kps[i, :counts[i]]
        assert (np.diff(k["octave"]) >= 0).all() and k["x"].min() >= 19 and k["x"].max() < 752 - 19 + 1
        assert counts[i] <= 1200 + 3 * 8


def test_kitti_stereo_pair_full_size():
    """BASELINE config 2: 1241 x 376 left/right, 2000 keypoints, stereo matching (also in test_gpu_parity; here the
    invariants at full size: disparity range, depth = mbf / disparity, left-right row consistency)."""
    from multiagent_orb_slam2_b200.extractor import compute_stereo_matches
    left, right = synth.stereo_pair("blocks", 1241, 376, 7)
    gL, gR = ORBextractor(2000, 1.2, 8, 20, 7), ORBextractor(2000, 1.2, 8, 20, 7)
    kL, _ = gL(left)
    kR, _ = gR(right)
    mbf = np.float32(386.1448)
    mb = np.float32(mbf / np.float32(718.856))
    u, d, kept = compute_stereo_matches(gL, gR, mbf, mb)
    u, d = u[:len(kL)], d[:len(kL)]
    ok = u >= 0
    assert kept == ok.sum() and kept > 300
    disp = kL["x"][ok] - u[ok]
    assert (disp >= 0).all() and (disp < mbf / mb).all()
    assert np.allclose(d[ok], mbf / np.maximum(disp, 0.01), rtol=1e-5)
    assert ((d >= 0) == ok).all()


def test_descriptor_mismatch_fraction_over_many_keypoints():
    """North-star clause: descriptor bits may differ only where a float rounds across a pattern-rotation boundary,
    with that fraction stated. The device rounds cos/sin of the angle correctly (double, rounded once); glibc's
    cosf/sinf used by the oracle differ from that by 1 ulp for ~1.3 % of angles, which moves a sample point for ~2e-4 of
    those keypoints => expected ~5e-6 differing descriptors. Measured here over > 90 k keypoints; angles and
    keypoints themselves must be identical everywhere."""
    imgs = np.stack([synth.image(kind, 640, 480, 500 + s) for s in range(32) for kind in ("blocks", "blurnoise", "noise")])
    g = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=len(imgs))
    kps, desc, counts = g.extract_batch(imgs)
    o = O.OracleExtractor()
    total = differing = 0
    for i in range(len(imgs)):
        ok, od = o(imgs[i])
        n = int(counts[i])
        assert n == len(ok)
        got = np.stack([kps[i, :n][f] for f in ("x", "y", "size", "angle", "response")], 1)
        assert np.array_equal(got.view(np.uint32), ok[:, :5].view(np.uint32))
        total += n
        differing += int((desc[i, :n] != od).any(1).sum())
    print("descriptor rows differing: %d of %d keypoints (%.2e)" % (differing, total, differing / total))
    assert total > 90_000 and differing / total < 1e-4
