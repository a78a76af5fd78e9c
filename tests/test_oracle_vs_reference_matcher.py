"""Pins the matcher oracle to the reference ITSELF (rows a-9 ... a-19, f-2 ... f-4 of SURVEY.md section 8): the reference's
own src/ORBmatcher.cc, Frame.cc, KeyFrame.cc, MapPoint.cc, Map.cc, KeyFrameDatabase.cc compiled unmodified into
oracle/_ref/libref_slam.so (oracle/build_ref.sh, oracle/ref_slam_wrap.cc) run on the same scenes as the scalar restatements
of oracle/oracle_lib.py. Integer results (match sets, candidate lists, grid cells): identical. Float outputs (projections,
viewing cosines, stereo coordinates): identical bit patterns."""
import numpy as np
import pytest

import guided_scenario as G
import oracle_lib as O
import ref_slam as R
from guided_reference import RefScene, fuse_graph_model, reference_results
from multiagent_orb_slam2_b200 import synth

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_slam.so not built (reference tree absent)")
f32 = np.float32
W, H = G.W, G.H


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_equals_reference_guided_searches(seed):
    views, shift, scale = G.extract_pair_cpu(seed)
    sc = G.make_scenario(seed, views, shift, scale)
    want = G.oracle_results(sc)
    got, fuse, tail, frustum = reference_results(sc)
    names = G.NAMES
    # searches 1-6 (a-10, a-11 x2, a-12 x2, a-13 KF-F, a-14 x2)
    for name, g, w in zip(names[:16], got, want[:16]):
        assert np.array_equal(g, w), (name, g[:20], w[:20])
    # Frame::isInFrustum == project_points mode 0 (bit patterns of u, v, ur, viewCos)
    fr = sc["frustum"]
    alive = fr["alive"].astype(bool)
    assert np.array_equal(frustum[:, 0] != 0, alive)
    for col, key in ((1, "u"), (2, "v"), (3, "ur"), (5, "view_cos")):
        assert np.array_equal(bits(frustum[alive, col]), bits(fr[key][alive])), key
    assert np.array_equal(frustum[alive, 4].astype(np.int32), fr["level"][alive])
    # search 8 and 9: Fuse(KF, Scw) and SearchBySim3
    for name, g, w in zip(names[19:], tail, want[19:]):
        assert np.array_equal(g, w), (name, g[:20], w[:20])
    # search 7, Fuse(KF, vpMapPoints): same candidate per map point (oracle) pushed through the reference's graph calls
    F1 = O.OracleFrame(sc["k"][1], sc["d"][1], W, H)
    F1.scale, F1.uright = scale, sc["uright"][1]
    T1 = sc["T"][1]
    c1 = dict(sc["cam"], Rcw=T1[:3, :3], tcw=T1[:3, 3], Ow=O.minus_Rt_t(T1[:3, :3], T1[:3, 3]))
    pr = O.project_points(c1, sc["pos"], sc["nrm"], sc["max_d"], sc["min_d"], 1, 0.5, float(G.TH[4]))
    model_state = None
    # the reference re-reads IsInKeyFrame / isBad while it runs, so the gate of later points sees earlier fusions
    in_kf1 = np.zeros(len(sc["pos"]), bool); in_kf1[sc["assoc"][1][sc["assoc"][1] >= 0]] = True
    sigma2 = (scale * scale).astype(f32)
    inv_sigma2 = (f32(1.0) / sigma2).astype(f32)
    qq = dict(valid=pr["alive"].astype(bool) & ~sc["bad"] & ~in_kf1, u=pr["u"], v=pr["v"], ur=pr["ur"], level=pr["level"], desc=sc["mdesc"])
    _, best, _ = O.fuse_kf_mappoints(F1, inv_sigma2, qq, float(G.TH[4]))
    model_state = fuse_graph_model(sc, best)
    assert fuse["n"] == model_state["n"]
    for key in ("held", "bad", "nobs", "replaced"):
        assert np.array_equal(np.asarray(fuse[key]).astype(np.int64), np.asarray(model_state[key]).astype(np.int64)), key
    # not vacuous
    counts = dict(zip(names, want))
    for key, least in [("a10 n", 100), ("a11 stereo n", 50), ("a11 mono n", 50), ("a12 cur-kf n", 50), ("a12 kf-scw n", 30), ("bow n", 50),
                       ("tri n", 10), ("fuse n", 30), ("fuse-scw n", 30), ("sim3 n", 10)]:
        assert counts[key][0] >= least, (key, counts[key][0])


# ---- Frame frontend: constructors on images, grid, GetFeaturesInArea, UndistortKeyPoints, stereo, RGB-D ---------------
TUM1 = dict(K=[517.306408, 516.469215, 318.643040, 255.313989], dist=[0.262383, -0.953104, -0.005358, 0.002628, 1.163314])  # TUM1.yaml
KITTI = dict(K=[718.856, 718.856, 607.1928, 185.2157], bf=386.1448)                                                        # KITTI00-02.yaml


def _keys6(fr):
    return fr["kps"]


@pytest.mark.parametrize("kind,seed", [("blocks", 3), ("blurnoise", 4)])
def test_mono_frame_constructor_grid_and_features_in_area(kind, seed):
    """Frame(mono) on an image (src/Frame.cc:174-228): keypoints / descriptors == oracle extractor, the 64 x 48 grid ==
    OracleFrame's, GetFeaturesInArea == features_in_area (same indices, same order) incl. the level filters."""
    img = synth.image(kind, W, H, seed)
    w = R.World()
    f = w.frame_images(0, img, TUM1["K"])
    got = w.frame_get(f)
    k, d = O.OracleExtractor(1000, 1.2, 8, 20, 7)(img)
    assert np.array_equal(bits(got["kps"]), bits(k)) and np.array_equal(got["desc"], d)
    assert np.array_equal(bits(got["un"]), bits(k[:, :2]))            # no distortion: mvKeysUn = mvKeys
    assert np.all(got["uright"] == -1) and np.all(got["depth"] == -1)
    F = O.OracleFrame(k, d, W, H)
    counts, items = w.frame_grid(f)
    want_items = [i for ix in range(64) for iy in range(48) for i in F.grid[ix][iy]]
    assert counts.tolist() == [[len(F.grid[ix][iy]) for iy in range(48)] for ix in range(64)]
    assert items.tolist() == want_items and len(want_items) > 900
    rng = np.random.default_rng(seed)
    nonempty = 0
    for _ in range(300):
        x, y = f32(rng.uniform(-30, W + 30)), f32(rng.uniform(-30, H + 30))
        r = f32(rng.choice([3.0, 7.5, 15.0, 40.0, 100.0, 700.0]))
        lo, hi = [(-1, -1), (0, 0), (2, 3), (0, 4), (3, -1), (-1, 2)][rng.integers(0, 6)]
        a = w.frame_features_in_area(f, x, y, r, lo, hi)
        assert a == F.features_in_area(x, y, r, lo, hi)
        nonempty += len(a) > 0
    assert nonempty > 100
    # KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:589-628): same cells, no level filter
    kf = w.keyframe(f)
    for _ in range(100):
        x, y, r = f32(rng.uniform(0, W)), f32(rng.uniform(0, H)), f32(rng.choice([4.0, 20.0, 90.0]))
        assert w.kf_features_in_area(kf, x, y, r) == F.features_in_area(x, y, r)


def test_undistort_keypoints_and_image_bounds_with_tum1_distortion():
    """Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:404-464) with the TUM1 distortion: the reference run on
    the cv2-pinned undistortPoints stand-in == oracle restatement == cv2 itself; the grid then uses the undistorted
    bounds."""
    cv2 = pytest.importorskip("cv2")
    img = synth.image("blocks", W, H, 5)
    w = R.World()
    f = w.frame_images(0, img, TUM1["K"], dist=TUM1["dist"])
    got = w.frame_get(f)
    k, d = O.OracleExtractor(1000, 1.2, 8, 20, 7)(img)
    K = np.array([[TUM1["K"][0], 0, TUM1["K"][2]], [0, TUM1["K"][1], TUM1["K"][3]], [0, 0, 1]], f32)
    dist = np.array(TUM1["dist"], f32)
    un = O.undistort_points(k[:, :2], K, dist)
    assert np.array_equal(bits(got["un"]), bits(un))
    assert np.array_equal(bits(un), bits(cv2.undistortPoints(k[:, :2].reshape(-1, 1, 2).copy(), K, dist, None, K).reshape(-1, 2)))
    assert np.abs(un - k[:, :2]).max() > 1.0                           # the distortion is not a no-op
    b = w.frame_bounds(f)
    want = O.image_bounds(W, H, K, dist)
    assert np.array_equal(bits(b[:4]), bits(want))
    Fo = O.OracleFrame(np.concatenate([un, k[:, 2:]], 1), d, W, H, bounds=want)
    counts, items = w.frame_grid(f)
    assert items.tolist() == [i for ix in range(64) for iy in range(48) for i in Fo.grid[ix][iy]]
    rng = np.random.default_rng(1)
    for _ in range(100):
        x, y, r = f32(rng.uniform(-20, W + 20)), f32(rng.uniform(-20, H + 20)), f32(rng.choice([5.0, 30.0, 120.0]))
        assert w.frame_features_in_area(f, x, y, r, 0, 3) == Fo.features_in_area(x, y, r, 0, 3)


@pytest.mark.parametrize("wd,ht,nf,kind,seed", [(1241, 376, 2000, "blocks", 0), (752, 480, 1200, "blocks", 1), (640, 480, 1000, "blurnoise", 2)])
def test_stereo_frame_constructor_equals_oracle(wd, ht, nf, kind, seed):
    """Frame(stereo) (src/Frame.cc:61-117): two extractor threads + ComputeStereoMatches (466-640). mvuRight / mvDepth bit
    patterns == oracle stereo_match (which the CUDA stereo kernel is tested against)."""
    left, right = synth.stereo_pair(kind, wd, ht, seed)
    mbf, fx = f32(KITTI["bf"]), f32(KITTI["K"][0])
    mb = f32(mbf / fx)
    w = R.World()
    f = w.frame_images(1, left, KITTI["K"], imgR=right, bf=float(mbf), nfeatures=nf, mb_before=float(mb))
    got = w.frame_get(f)
    oL, oR = O.OracleExtractor(nf, 1.2, 8, 20, 7), O.OracleExtractor(nf, 1.2, 8, 20, 7)
    kL, dL = oL(left)
    oR(right)
    ou, od, kept = O.stereo_match(oL, oR, mbf, mb)
    assert np.array_equal(bits(got["kps"]), bits(kL)) and np.array_equal(got["desc"], dL)
    assert np.array_equal(bits(got["uright"]), bits(ou)) and np.array_equal(bits(got["depth"]), bits(od))
    assert kept == int((got["depth"] > 0).sum()) and kept > 100


def test_rgbd_frame_constructor_equals_restatement():
    """Frame(RGB-D) (src/Frame.cc:119-172) -> ComputeStereoFromRGBD (643-664), with distortion so that kp and kpU differ."""
    rng = np.random.default_rng(4)
    img = synth.image("blocks", W, H, 90)
    depth = rng.uniform(0.3, 8.0, (H, W)).astype(f32)
    depth[rng.random((H, W)) < 0.2] = 0.0
    depth[rng.random((H, W)) < 0.02] = -1.0
    for dist in ([0, 0, 0, 0], TUM1["dist"]):
        w = R.World()
        f = w.frame_images(2, img, TUM1["K"], depth=depth, dist=dist, bf=40.0)
        got = w.frame_get(f)
        ur, dz = O.stereo_from_rgbd(got["kps"], got["un"], depth, 40.0)
        assert np.array_equal(bits(got["uright"]), bits(ur)) and np.array_equal(bits(got["depth"]), bits(dz))
        assert 0.6 < (dz > 0).mean() < 0.9


# ---- the two searches that are not part of the guided scenario -------------------------------------------------------
@pytest.mark.parametrize("seed,window,ratio,ori", [(0, 100, 0.9, True), (1, 30, 0.9, True), (2, 100, 0.7, False), (3, 10, 0.9, True)])
def test_oracle_equals_reference_search_for_initialization(seed, window, ratio, ori):
    views, _, _ = G.extract_pair_cpu(seed + 20)
    (k0, d0), (k1, d1) = views
    w = R.World()
    f0, f1 = (w.frame_arrays(k, d, TUM1["K"], W, H) for k, d in views)
    F0, F1 = O.OracleFrame(k0, d0, W, H), O.OracleFrame(k1, d1, W, H)
    prev = k0[:, :2].copy()
    for _ in range(2):   # the second call starts from the updated vbPrevMatched, like consecutive initialisation attempts
        want_prev = prev.copy()
        wn, wm = O.search_for_initialization(F0, F1, want_prev, window, ratio, ori)
        gn, gm, got_prev = w.search_for_initialization(f0, f1, prev, window, ratio, ori)
        assert gn == wn and np.array_equal(gm, wm) and np.array_equal(bits(got_prev), bits(want_prev))
        prev = want_prev
    assert wn > (20 if window >= 30 else 3)


@pytest.mark.parametrize("seed,ratio,ori", [(0, 0.75, True), (1, 0.75, False), (2, 0.9, True)])
def test_oracle_equals_reference_search_by_bow_kf_kf(seed, ratio, ori):
    """SearchByBoW(KF1, KF2) (src/ORBmatcher.cc:524-657), the matcher of MapFusion::ComputeSim3 / CovisibilityDiscovery
    (src/MapFusion.cc:275, 849)."""
    views, shift, scale = G.extract_pair_cpu(seed + 30)
    sc = G.make_scenario(seed + 30, views, shift, scale)
    s = RefScene(sc); s.hold(0); s.hold(1)
    assoc0, assoc1 = sc["assoc"]
    gn, gm = s.w.search_by_bow_kf_kf(s.kf[0], s.kf[1], len(assoc0), ratio, ori)
    v0 = (assoc0 >= 0) & ~sc["bad"][np.maximum(assoc0, 0)]
    v1 = (assoc1 >= 0) & ~sc["bad"][np.maximum(assoc1, 0)]
    wn, wm = O.search_by_bow_kf_kf(sc["d"][0], sc["fv"][0], v0, sc["k"][0][:, 3], sc["d"][1], sc["fv"][1], v1, sc["k"][1][:, 3], ratio, ori)
    assert gn == wn and np.array_equal(gm, np.where(wm >= 0, assoc1[np.maximum(wm, 0)], -1))
    assert wn >= 10


def test_oracle_equals_reference_distinctive_descriptor():
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:246-311). The reference walks its observations in
    std::map<KeyFrame*, size_t> order, i.e. by heap address; ties between medians are therefore allocator dependent, so
    the check is: the chosen descriptor attains the least median, and equals the oracle's choice when that is unique."""
    rng = np.random.default_rng(7)
    w = R.World()
    kfs, descs = [], []
    base = synth.descriptors(1, 3)[0]
    for i in range(12):
        n = 3
        k = np.zeros((n, 6), f32); k[:, 0] = 100 + 10 * np.arange(n); k[:, 1] = 100; k[:, 2] = 31
        b = np.unpackbits(np.tile(base, (n, 1)), axis=1)
        for r in range(n):
            b[r, rng.choice(256, rng.integers(0, 60), replace=False)] ^= 1
        d = np.packbits(b, axis=1)
        kfs.append(w.keyframe(w.frame_arrays(k, d, TUM1["K"], W, H)))
        descs.append(d)
    unique = 0
    for n_obs in (1, 2, 3, 4, 7, 12):
        m = w.mappoint([0, 0, 5], kfs[0])
        rows = []
        for j in range(n_obs):
            idx = int(rng.integers(0, 3))
            w.observe(m, kfs[j], idx)
            rows.append(descs[j][idx])
        w.mp_compute_distinctive(m)
        got = w.mp_get(m)["desc"]
        rows = np.array(rows)
        med = [sorted(0 if a == b2 else O.hamming(rows[a], rows[b2]) for b2 in range(n_obs))[int(0.5 * (n_obs - 1))] for a in range(n_obs)]
        winners = [a for a in range(n_obs) if med[a] == min(med)]
        assert any(np.array_equal(got, rows[a]) for a in winners)
        if len(winners) == 1:
            assert np.array_equal(got, rows[O.distinctive_descriptor(rows)])
            unique += 1
    assert unique >= 3


# ---- KeyFrameDatabase ----------------------------------------------------------------------------
@pytest.mark.parametrize("seed,n_kf", [(0, 400), (1, 120)])
def test_oracle_equals_reference_keyframe_database(tmp_path, seed, n_kf):
    """KeyFrameDatabase::add / erase / DetectLoopCandidates (src/KeyFrameDatabase.cc:40-197) on the reference's own
    KeyFrame covisibility graph, against oracle_lib.detect_loop_candidates (which the CUDA database is tested against)."""
    from test_kfdb import bow_vectors
    rng = np.random.default_rng(seed)
    voc = synth.vocabulary(10, 4, seed=3)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, voc)
    w = R.World(path)
    n_words = 10 ** 4
    vecs, _ = bow_vectors(rng, n_kf + 6, n_words)
    k1 = np.array([[50, 50, 31, 0, 20, 0]], f32); d1 = np.zeros((1, 32), np.uint8)
    kfs = []
    for ids, wt in vecs:
        kf = w.keyframe(w.frame_arrays(k1, d1, TUM1["K"], W, H))
        w.kf_set_bowvec(kf, ids, wt)
        kfs.append(kf)
    for kf in kfs[:n_kf]:
        w.db_add(kf)
    alive = np.ones(n_kf, bool)
    for slot in rng.choice(n_kf, n_kf // 20, replace=False):
        w.db_erase(int(slot)); alive[slot] = False
    neigh = {}
    for s in range(n_kf):
        nb = [int(x) for x in rng.choice(np.delete(np.arange(n_kf), s), 10, replace=False)]
        neigh[s] = nb
        for rank, b in enumerate(nb):
            w.kf_add_connection(s, b, 100 - rank)          # distinct weights: GetBestCovisibilityKeyFrames order = nb
    # the fork's covisibility query and the relocalisation query read per-keyframe scores that they do not (always) write:
    # start every keyframe from a known value, as an explicit input of both sides
    covis_state = rng.uniform(0.0, 0.05, n_kf).astype(f32)
    reloc_state = rng.uniform(0.0, 0.05, n_kf).astype(f32)
    for s in range(n_kf):
        w.kf_set_scores(s, float(covis_state[s]), float(reloc_state[s]))
    hits = 0
    for qi in range(n_kf, n_kf + 6):
        q_ids, q_w = vecs[qi]
        connected = set(int(x) for x in rng.choice(n_kf, 15, replace=False))
        for c in connected:
            w.kf_add_connection(qi, c, 20)
        scores = sorted((O.l1_score(q_ids, q_w, *vecs[s]) for s in range(n_kf)), reverse=True)
        min_score = 0.5 * float(scores[max(3, n_kf // 50)])
        want = O.detect_loop_candidates(vecs[:n_kf], alive, q_ids, q_w, min_score, connected, lambda s: neigh[s])
        got = w.db_detect_loop_candidates(qi, min_score)
        assert got == want
        hits += len(want) > 0
        # DetectCovisibilityCandidates (199-308): ignore list, stale mCovisScore
        ignore = [int(x) for x in rng.choice(n_kf, 12, replace=False)]
        want_c = O.detect_covisibility_candidates(vecs[:n_kf], alive, q_ids, q_w, min_score, ignore, lambda s: neigh[s], covis_state)
        got_c = w.db_detect_covisibility_candidates(qi, min_score, ignore)
        assert got_c == want_c and len(want_c) >= 1
        # DetectRelocalizationCandidates (310-420): a Frame with the query's BowVector; mRelocScore carries over between queries
        fq = w.frame_arrays(k1, d1, TUM1["K"], W, H)
        w.frame_set_bowvec(fq, q_ids, q_w)
        want_r = O.detect_relocalization_candidates(vecs[:n_kf], alive, q_ids, q_w, lambda s: neigh[s], reloc_state)
        got_r = w.db_detect_relocalization_candidates(fq)
        assert got_r == want_r and len(want_r) >= 1
        for s in range(0, n_kf, 9):
            assert np.float32(w.kf_reloc_score(s)) == reloc_state[s]
    assert hits >= 4
