"""Rows a-11 ... a-15 of SURVEY.md section 8: the projection / BoW guided searches of ORBmatcher through the product
(device window lists / list distances + ordered host replay, multiagent_orb_slam2_b200/guided.py) against the scalar
restatements in oracle/oracle_lib.py. Match sets must be identical (integer work: no tolerance)."""
import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import ORBextractor, ORBmatcher, synth

pytestmark = pytest.mark.gpu
W, H = 640, 480
f32 = np.float32


class Scene:
    """Two views of one synthetic scene; 'map points' are the keypoints of view A, projected into view B by the
    known shift plus noise, with a predicted level near their octave."""

    def __init__(self, seed, kind="blocks"):
        from multiagent_orb_slam2_b200.device_grid import DeviceFrameGrid
        a, b, (dx, dy) = synth.shifted_pair(kind, W, H, seed + 40)
        self.ex_a, self.ex_b = ORBextractor(1000, 1.2, 8, 20, 7), ORBextractor(1000, 1.2, 8, 20, 7)
        self.ka, self.da = self.ex_a(a)
        self.kb, self.db = self.ex_b(b)
        self.grid_a, self.grid_b = DeviceFrameGrid(self.ex_a), DeviceFrameGrid(self.ex_b)
        self.rng = rng = np.random.default_rng(seed)
        self.scale = self.ex_b.GetScaleFactors()
        self.sigma2 = self.ex_b.GetScaleSigmaSquares()
        self.inv_sigma2 = self.ex_b.GetInverseScaleSigmaSquares()
        self.Fa, self.Fb = self.oracle_frame(self.ka, self.da), self.oracle_frame(self.kb, self.db)
        self.dx, self.dy = dx, dy
        n = len(self.ka)
        self.q_ab = self.queries(self.ka, self.da, -dx, -dy, n)
        self.q_ba = self.queries(self.kb, self.db, dx, dy, len(self.kb))

    def oracle_frame(self, k, d):
        ok = np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"].astype(f32)], 1)
        F = O.OracleFrame(ok, d, W, H)
        F.scale = self.scale
        return F

    def queries(self, k, d, sx, sy, n):
        rng = self.rng
        u = (k["x"] + f32(sx) + rng.normal(0, 1.0, n)).astype(f32)
        v = (k["y"] + f32(sy) + rng.normal(0, 1.0, n)).astype(f32)
        lvl = np.clip(k["octave"].astype(np.int32) + rng.integers(-1, 2, n) * (rng.random(n) < 0.3), 0, 7).astype(np.int32)
        valid = (rng.random(n) < 0.9) & (u >= 0) & (u < W) & (v >= 0) & (v < H)
        return dict(valid=valid, u=u, v=v, ur=(u - rng.uniform(2, 40, n)).astype(f32), level=lvl, octave=k["octave"].astype(np.int32),
                    angle=k["angle"].astype(f32), desc=d, has_obs=rng.random(n) < 0.95)


@pytest.mark.parametrize("seed,th,mode", [(0, 15.0, "none"), (1, 7.0, "forward"), (2, 15.0, "backward"), (3, 30.0, "none")])
def test_search_by_projection_cur_last(seed, th, mode):
    s = Scene(seed)
    rng = s.rng
    uright = np.where(rng.random(len(s.kb)) < 0.5, s.kb["x"] - rng.uniform(2, 40, len(s.kb)), -1).astype(f32)
    occupied = rng.random(len(s.kb)) < 0.1
    s.Fb.uright, s.Fb.occupied = uright, occupied
    fw, bw = mode == "forward", mode == "backward"
    onm, oas = O.search_by_projection_cur_last(s.Fb, s.q_ab, th, fw, bw)
    gnm, gas = ORBmatcher(0.9, True).SearchByProjection_Cur_Last(s.grid_b, s.kb["angle"], uright, occupied, s.scale, s.q_ab, th, fw, bw)
    assert gnm == onm and np.array_equal(gas, oas)
    assert onm > 100


@pytest.mark.parametrize("seed,th,orb_dist", [(0, 10.0, 100), (1, 3.0, 64)])
def test_search_by_projection_cur_kf(seed, th, orb_dist):
    s = Scene(seed + 10)
    occupied = s.rng.random(len(s.kb)) < 0.2
    s.Fb.occupied = occupied
    onm, oas = O.search_by_projection_cur_kf(s.Fb, s.q_ab, th, orb_dist)
    gnm, gas = ORBmatcher(0.9, True).SearchByProjection_Cur_KF(s.grid_b, s.kb["angle"], occupied, s.scale, s.q_ab, th, orb_dist)
    assert gnm == onm and np.array_equal(gas, oas)
    assert onm > 100


@pytest.mark.parametrize("seed,th", [(0, 10), (1, 4)])
def test_search_by_projection_kf_sim3(seed, th):
    s = Scene(seed + 20)
    matched = s.rng.random(len(s.kb)) < 0.2
    onm, oas = O.search_by_projection_kf_sim3(s.Fb, matched, s.q_ab, th)
    gnm, gas = ORBmatcher(0.75, True).SearchByProjection_KF_Sim3(s.grid_b, matched, s.scale, s.q_ab, th)
    assert gnm == onm and np.array_equal(gas, oas)
    assert onm > 50


@pytest.mark.parametrize("seed,th", [(0, 3.0), (1, 6.0)])
def test_fuse_kf_mappoints(seed, th):
    s = Scene(seed + 30)
    uright = np.where(s.rng.random(len(s.kb)) < 0.5, s.kb["x"] - s.rng.uniform(2, 40, len(s.kb)), -1).astype(f32)
    # make some stereo candidates consistent with the queries' ur so that both branches of the chi2 gate accept
    s.Fb.uright = uright
    on, oi, od = O.fuse_kf_mappoints(s.Fb, s.inv_sigma2, s.q_ab, th)
    gn, gi, gd = ORBmatcher().Fuse_KF_MapPoints(s.grid_b, s.kb["x"], s.kb["y"], s.kb["octave"], uright, s.inv_sigma2, s.scale, s.q_ab, th)
    assert gn == on and np.array_equal(gi, oi) and np.array_equal(gd, od)
    assert on > 50


@pytest.mark.parametrize("seed,th", [(0, 4.0), (1, 7.5)])
def test_fuse_sim3_and_search_by_sim3(seed, th):
    s = Scene(seed + 40)
    m = ORBmatcher(0.75, True)
    oi = O.best_in_window(s.Fb, s.q_ab, th, 50)
    gn, gi, _ = m.Fuse_KF_Sim3(s.grid_b, s.scale, s.q_ab, th)
    assert np.array_equal(gi, oi) and gn == (oi >= 0).sum() and gn > 50
    on, om = O.search_by_sim3(s.Fa, s.Fb, s.q_ab, s.q_ba, th)
    gn2, gm = m.SearchBySim3(s.grid_a, s.grid_b, s.scale, s.scale, s.q_ab, s.q_ba, th)
    assert gn2 == on and np.array_equal(gm, om)
    assert on > 50


def _bow_case(seed, n1=900, n2=850, nodes=60):
    rng = np.random.default_rng(seed)
    d1 = synth.descriptors(n1, 10 + seed)
    twin = rng.integers(0, n1, n2)
    bits = np.unpackbits(d1[twin], axis=1)
    for i in range(n2):
        bits[i, rng.choice(256, rng.integers(0, 70), replace=False)] ^= 1
    d2 = np.packbits(bits, axis=1)
    node1 = rng.integers(0, nodes, n1)
    node2 = np.where(rng.random(n2) < 0.9, node1[twin], rng.integers(0, nodes, n2))
    fv = []
    for node in (node1, node2):
        d = {}
        for i, k in enumerate(node):
            d.setdefault(int(k), []).append(i)
        fv.append(sorted(d.items()))
    a1 = rng.uniform(0, 360, n1).astype(f32)
    a2 = ((a1[twin] + rng.normal(0, 8, n2)) % 360).astype(f32)
    return rng, d1, d2, fv[0], fv[1], a1, a2, twin


@pytest.mark.parametrize("seed,ratio", [(0, 0.7), (1, 0.9)])
def test_search_by_bow_kf_frame(seed, ratio):
    rng, d1, d2, fv1, fv2, a1, a2, _ = _bow_case(seed)
    v1 = rng.random(len(d1)) < 0.85
    onm, oas = O.search_by_bow_kf_f(d1, fv1, v1, a1, d2, fv2, a2, ratio)
    gnm, gas = ORBmatcher(ratio, True).SearchByBoW_KF_F(d1, fv1, v1, a1, d2, fv2, a2)
    assert gnm == onm and np.array_equal(gas, oas)
    assert onm > 100


@pytest.mark.parametrize("seed,only_stereo", [(0, False), (1, True), (2, False)])
def test_search_for_triangulation(seed, only_stereo):
    rng, d1, d2, fv1, fv2, a1, a2, twin = _bow_case(seed + 5, 700, 700)
    n1, n2 = len(d1), len(d2)
    # pure sideways translation between the two keyframes: F12 = [t]x up to scale, epipolar lines are image rows
    x1, y1 = rng.uniform(20, 620, n1).astype(f32), rng.uniform(20, 460, n1).astype(f32)
    x2 = (x1[twin] - rng.uniform(1, 60, n2)).astype(f32)
    y2 = (y1[twin] + rng.normal(0, 1.2, n2)).astype(f32)
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], f32) * f32(0.37)
    scale = (f32(1.2) ** np.arange(8)).astype(f32)
    sigma2 = (scale * scale).astype(f32)
    kf1 = dict(desc=d1, featvec=fv1, has_mp=rng.random(n1) < 0.3, uright=np.where(rng.random(n1) < 0.5, x1 - 5, -1).astype(f32),
               x=x1, y=y1, angle=a1, octave=rng.integers(0, 8, n1))
    kf2 = dict(desc=d2, featvec=fv2, has_mp=rng.random(n2) < 0.3, uright=np.where(rng.random(n2) < 0.5, x2 - 5, -1).astype(f32),
               x=x2, y=y2, angle=a2, octave=rng.integers(0, 8, n2))
    epipole = (f32(300.0), f32(240.0))  # inside the image so that the epipole-distance gate fires for some candidates
    check_ori = seed != 2
    want = O.search_for_triangulation(kf1, kf2, F12, epipole, scale, sigma2, only_stereo, check_ori=check_ori)
    got = ORBmatcher(0.6, check_ori).SearchForTriangulation(kf1, kf2, F12, epipole, scale, sigma2, only_stereo)
    assert got == want
    assert len(want) > 30
