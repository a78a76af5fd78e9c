"""The drop-in boundary exercised from C++ (include/orbslam2_b200/*.h through liborb_b200.so) and from
the Python mirror, against the oracle: ORBextractor::operator(), mvImagePyramid,
ORBmatcher::SearchForInitialization (stateful, windowed) and the brute-force ratio search."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import ORBextractor, ORBmatcher, synth
from multiagent_orb_slam2_b200.matcher import FrameGrid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_facade_test(tmp):
    exe = os.path.join(tmp, "facade_test")
    lib = os.path.join(ROOT, "multiagent_orb_slam2_b200", "lib")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_test.cc"),
                           "-o", exe, "-L" + lib, "-lorb_b200", "-Wl,-rpath," + lib])
    return exe


def test_facade_headers_compile(tmp_path):
    """CPU: the facade compiles and links against the C-ABI library with the OpenCV stand-in types."""
    assert os.path.exists(build_facade_test(str(tmp_path)))


def oracle_init_search(a, b, nf, window):
    o = O.OracleExtractor(2 * nf, 1.2, 8, 20, 7)
    ka, da = o(a)
    lvl3_a = o.level(3)["img"].copy()   # the oracle object holds frame b's pyramid after the next call
    kb, db = o(b)
    F1, F2 = O.OracleFrame(ka, da, a.shape[1], a.shape[0]), O.OracleFrame(kb, db, a.shape[1], a.shape[0])
    prev = ka[:, :2].copy()
    nm, m12 = O.search_for_initialization(F1, F2, prev, window, 0.9, True)
    return lvl3_a, ka, da, kb, db, nm, m12, prev


@pytest.mark.gpu
def test_cpp_facade_matches_oracle(tmp_path):
    exe = build_facade_test(str(tmp_path))
    w, h, nf = 640, 480, 500
    a, b, _ = synth.shifted_pair("blocks", w, h, 3)
    pa, pb, out = [str(tmp_path / n) for n in ("a.raw", "b.raw", "out.bin")]
    a.tofile(pa); b.tofile(pb)
    print(subprocess.check_output([exe, str(w), str(h), pa, pb, out, str(nf)], text=True))
    lvl3_a, ka, da, kb, db, nm, m12, _ = oracle_init_search(a, b, nf, 100)
    buf = open(out, "rb").read()
    pos = 0

    def take(dtype, count):
        nonlocal pos
        arr = np.frombuffer(buf, dtype, count, pos)
        pos += arr.nbytes
        return arr

    for k, (ok, od) in enumerate([(ka, da), (kb, db)]):
        n = int(take(np.int32, 1)[0])
        assert n == len(ok)
        assert np.array_equal(take(np.float32, 6 * n).reshape(n, 6).view(np.uint32), ok.view(np.uint32))
        assert np.array_equal(take(np.uint8, 32 * n).reshape(n, 32), od)
        if k == 0:
            pw, ph = take(np.int32, 2)
            assert (ph, pw) == lvl3_a.shape
            assert np.array_equal(take(np.uint8, pw * ph).reshape(ph, pw), lvl3_a)   # mvImagePyramid[3] of frame a through the facade
    assert int(take(np.int32, 1)[0]) == nm
    n1 = int(take(np.int32, 1)[0])
    assert np.array_equal(take(np.int32, n1), m12)
    assert nm > 50
    nb = int(take(np.int32, 1)[0])
    bf = take(np.int32, n1)
    oi, o1, o2 = O.knn2(da, db)
    ref = np.where((o1 < 50) & (o1.astype(np.float32) < np.float32(0.9) * o2.astype(np.float32)), oi, -1)
    assert np.array_equal(bf, ref) and nb == int((ref >= 0).sum())
    # SearchByBoW(KF,KF) of the C++ facade vs the oracle, with the driver's synthetic feature vectors / map points
    nbow = int(take(np.int32, 1)[0])
    m12bow = take(np.int32, n1)

    def kf(kps):
        fv = {}
        for i in range(len(kps)):
            fv.setdefault((int(kps[i, 0] / np.float32(80)) + 8 * int(kps[i, 1] / np.float32(96))) % 40, []).append(i)
        idx = np.arange(len(kps))
        return sorted(fv.items()), (idx % 8 != 3) & ~(idx % 13 == 5)

    fv1, v1 = kf(ka)
    fv2, v2 = kf(kb)
    onm, om = O.search_by_bow_kf_kf(da, fv1, v1, ka[:, 3], db, fv2, v2, kb[:, 3], 0.75)
    assert nbow == onm and np.array_equal(m12bow, om) and onm > 20


@pytest.mark.gpu
@pytest.mark.parametrize("seed,window", [(0, 100), (1, 30)])
def test_python_search_for_initialization_matches_oracle(seed, window):
    w, h, nf = 640, 480, 500
    a, b, _ = synth.shifted_pair("blocks", w, h, seed)
    _, ka, da, kb, db, nm, m12, prev_ref = oracle_init_search(a, b, nf, window)
    ex, ex_b = ORBextractor(2 * nf, 1.2, 8, 20, 7), ORBextractor(2 * nf, 1.2, 8, 20, 7)
    ga, gda = ex(a)
    gb, gdb = ex_b(b)
    F1, F2 = FrameGrid(ga, gda, w, h), FrameGrid(gb, gdb, w, h)
    prev = np.stack([ga["x"], ga["y"]], 1).astype(np.float32)
    gnm, gm12 = ORBmatcher(0.9, True).SearchForInitialization(F1, F2, prev, window)
    assert gnm == nm and np.array_equal(gm12, m12)
    assert np.array_equal(prev.view(np.uint32), prev_ref.view(np.uint32))
    assert nm > 30
    # same search with the candidate gate on the device (Frame grid + window lists in HBM)
    from multiagent_orb_slam2_b200.device_grid import DeviceFrameGrid
    prev2 = np.stack([ga["x"], ga["y"]], 1).astype(np.float32)
    dnm, dm12 = ORBmatcher(0.9, True).SearchForInitialization(F1, F2, prev2, window, device_grid=DeviceFrameGrid(ex_b))
    assert dnm == nm and np.array_equal(dm12, m12) and np.array_equal(prev2.view(np.uint32), prev_ref.view(np.uint32))


def _feature_vectors(rng, n1, n2, twin_of_2, nodes=60):
    """Synthetic DBoW2 FeatureVectors: every feature gets a vocabulary node; a feature of frame 2 that is a
    noisy copy of feature t of frame 1 lands in t's node 90 % of the time."""
    node1 = rng.integers(0, nodes, n1)
    node2 = np.where(rng.random(n2) < 0.9, node1[twin_of_2], rng.integers(0, nodes, n2))
    fv = []
    for node in (node1, node2):
        d = {}
        for i, k in enumerate(node):
            d.setdefault(int(k), []).append(i)
        fv.append(sorted(d.items()))
    return fv


@pytest.mark.gpu
@pytest.mark.parametrize("seed,ratio", [(0, 0.75), (1, 0.9)])
def test_search_by_bow_kf_kf_matches_oracle(seed, ratio):
    """The matcher MapFusion calls (src/MapFusion.cc:275): SearchByBoW(KF,KF) with its vbMatched2 state."""
    rng = np.random.default_rng(seed)
    n1, n2 = 900, 850
    d1 = synth.descriptors(n1, 10 + seed)
    twin = rng.integers(0, n1, n2)
    d2 = d1[twin].copy()
    bits = np.unpackbits(d2, axis=1)
    for i in range(n2):
        bits[i, rng.choice(256, rng.integers(0, 70), replace=False)] ^= 1
    d2 = np.packbits(bits, axis=1)
    fv1, fv2 = _feature_vectors(rng, n1, n2, twin)
    v1, v2 = rng.random(n1) < 0.85, rng.random(n2) < 0.85
    a1 = rng.uniform(0, 360, n1).astype(np.float32)
    a2 = (a1[twin] + rng.normal(0, 8, n2)).astype(np.float32) % np.float32(360)
    onm, om = O.search_by_bow_kf_kf(d1, fv1, v1, a1, d2, fv2, v2, a2, ratio)
    gnm, gm = ORBmatcher(ratio, True).SearchByBoW_KF_KF(d1, fv1, v1, a1, d2, fv2, v2, a2)
    assert gnm == onm and np.array_equal(gm, om)
    assert (om >= 0).sum() > 100


@pytest.mark.gpu
def test_distinctive_descriptor_matches_oracle():
    rng = np.random.default_rng(3)
    m = ORBmatcher()
    for n in (1, 2, 3, 8, 30):
        base = synth.descriptors(1, 50 + n)
        obs = synth.descriptors(n, 60 + n, dup_from=base, max_flips=40)
        assert m.ComputeDistinctiveDescriptor(obs) == O.distinctive_descriptor(obs)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,radius", [(0, 15.0), (1, 40.0), (2, 3.0)])
def test_device_grid_window_search_matches_oracle(seed, radius):
    """GetFeaturesInArea + best/second loop entirely on the device vs the oracle grid walk (same candidate order,
    so the first-minimum index and the octaves of best / second agree even with distance ties)."""
    from multiagent_orb_slam2_b200.device_grid import DeviceFrameGrid
    w, h = 640, 480
    a, b, (dx, dy) = synth.shifted_pair("blocks", w, h, seed)
    ex_a, ex_b = ORBextractor(1000, 1.2, 8, 20, 7), ORBextractor(1000, 1.2, 8, 20, 7)
    ka, da = ex_a(a)
    kb, db = ex_b(b)
    grid = DeviceFrameGrid(ex_b)
    # queries: keypoints of frame a projected by the known shift; level window like SearchByProjection (pred-1 .. pred)
    qx, qy = ka["x"] - np.float32(dx), ka["y"] - np.float32(dy)
    lo = np.maximum(ka["octave"] - 1, 0).astype(np.int32) * (seed != 2) + (-1) * (seed == 2)
    hi = ka["octave"].astype(np.int32) * (seed != 2) + (-1) * (seed == 2)
    r = (np.float32(radius) * ex_a.GetScaleFactors()[ka["octave"]]).astype(np.float32)
    gi, g1, g2, gl1, gl2 = grid.window_knn2(da, qx, qy, r, lo, hi)
    okb = np.stack([kb["x"], kb["y"], kb["size"], kb["angle"], kb["response"], kb["octave"].astype(np.float32)], 1)
    F = O.OracleFrame(okb, db, w, h)
    hits = 0
    for i in range(len(ka)):
        cand = F.features_in_area(qx[i], qy[i], r[i], int(lo[i]), int(hi[i]))
        b1 = b2 = 256
        bi = l1 = l2 = -1
        for j in cand:
            d = O.hamming(da[i], db[j])
            if d < b1:
                b2, l2, b1, l1, bi = b1, l1, d, int(kb["octave"][j]), j
            elif d < b2:
                b2, l2 = d, int(kb["octave"][j])
        assert (gi[i], g1[i], g2[i], gl1[i], gl2[i]) == (bi, b1, b2, l1, l2), i
        hits += bi >= 0
    assert hits > len(ka) // 3


@pytest.mark.gpu
@pytest.mark.parametrize("seed,th", [(0, 1.0), (1, 3.0), (2, 5.0)])
def test_search_by_projection_frame_mappoints_matches_oracle(seed, th):
    """SearchByProjection(Frame, MapPoints, th) (src/ORBmatcher.cc:45-131): local map points = the keypoints of a shifted
    view, projected into the frame by the known shift with a little noise; device window lists + host replay vs oracle."""
    from multiagent_orb_slam2_b200.device_grid import DeviceFrameGrid
    w, h = 640, 480
    a, b, (dx, dy) = synth.shifted_pair("blocks", w, h, seed + 20)
    ex_a, ex_b = ORBextractor(1000, 1.2, 8, 20, 7), ORBextractor(1000, 1.2, 8, 20, 7)
    ka, da = ex_a(a)
    kb, db = ex_b(b)
    rng = np.random.default_rng(seed)
    n = len(ka)
    mp = dict(in_view=rng.random(n) < 0.9, bad=rng.random(n) < 0.05, level=ka["octave"].astype(np.int32),
              view_cos=rng.uniform(0.99, 1.0, n).astype(np.float32),
              proj_x=(ka["x"] - np.float32(dx) + rng.normal(0, 1.0, n)).astype(np.float32),
              proj_y=(ka["y"] - np.float32(dy) + rng.normal(0, 1.0, n)).astype(np.float32), desc=da)
    mp["proj_xr"] = (mp["proj_x"] - rng.uniform(2, 40, n)).astype(np.float32)
    uright = np.where(rng.random(len(kb)) < 0.5, kb["x"] - rng.uniform(2, 40, len(kb)), -1).astype(np.float32)
    occupied = rng.random(len(kb)) < 0.1
    scale = ex_b.GetScaleFactors()
    okb = np.stack([kb["x"], kb["y"], kb["size"], kb["angle"], kb["response"], kb["octave"].astype(np.float32)], 1)
    F = O.OracleFrame(okb, db, w, h)
    F.scale, F.uright, F.occupied = scale, uright, occupied
    onm, oas = O.search_by_projection_frame_mappoints(F, mp, th, 0.8)
    gnm, gas = ORBmatcher(0.8).SearchByProjection_Frame_MapPoints(DeviceFrameGrid(ex_b), kb["octave"], uright, occupied, scale, mp, th)
    assert gnm == onm and np.array_equal(gas, oas)
    assert onm > 100


@pytest.mark.gpu
def test_device_grid_built_and_queried_on_a_side_stream_right_after_process_async():
    """The grid build and the window queries run on torch's current stream (never the legacy NULL stream, which the
    library's non-blocking streams are not ordered against): built under a side stream immediately after an asynchronous
    frontend launch - no host synchronisation in between - the candidate lists must still be those of the finished
    extraction; queries from a third stream wait for the build through its event."""
    import torch
    from multiagent_orb_slam2_b200.device_grid import DeviceFrameGrid
    from multiagent_orb_slam2_b200.frontend import AgentFrontend
    w, h = 640, 480
    frames = []
    for s in range(4):
        a, b, _ = synth.shifted_pair("blocks", w, h, s + 60)
        frames += [a, b]
    imgs = torch.from_numpy(np.stack(frames)).pin_memory()
    fe = AgentFrontend(w, h, max_batch=8)
    o = O.OracleExtractor()
    side, third = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(3):   # repeated so that a missing dependency would race against a busy device
        with torch.cuda.stream(side):
            fe.process_async(imgs)
            grid = DeviceFrameGrid(fe.ex, frame=5)
        k5, d5 = o(frames[5])
        k4, d4 = o(frames[4])
        qx, qy = k4[:, 0], k4[:, 1]
        r = np.full(len(k4), 20.0, np.float32)
        with torch.cuda.stream(third):
            gi, g1, g2, _, _ = grid.window_knn2(d4, qx, qy, r)
        F = O.OracleFrame(k5, d5, w, h)
        for i in range(0, len(k4), 7):
            cand = F.features_in_area(qx[i], qy[i], r[i])
            ds = [O.hamming(d4[i], d5[j]) for j in cand]
            want = (cand[int(np.argmin(ds))], min(ds)) if cand else (-1, 256)
            assert (gi[i], g1[i]) == want, (rep, i)
        grid.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,radius,levels", [(0, 100.0, (0, 0)), (1, 20.0, (-1, -1)), (2, 45.0, (1, 3))])
def test_host_buffer_window_lists_match_oracle(seed, radius, levels):
    """orbm_window_lists - the facade's device-side gate for host-resident frames (SearchForInitialization on the reference's
    Frame): grid rebuilt on the device from host keypoints, candidates in GetFeaturesInArea order with their distances;
    also the ORB_ECAPACITY protocol (needed room reported, second call fits) and windows of radius 0."""
    import ctypes as C
    from multiagent_orb_slam2_b200 import _lib
    from multiagent_orb_slam2_b200.extractor import KP_DTYPE
    L = _lib.lib()
    w, h = 640, 480
    a, b, (dx, dy) = synth.shifted_pair("blocks", w, h, seed + 40)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    ka, da = ex(a)
    kb, db = ex(b)
    ka, da, kb, db = ka.copy(), da.copy(), kb.copy(), db.copy()
    qx, qy = (ka["x"] - np.float32(dx)).astype(np.float32), (ka["y"] - np.float32(dy)).astype(np.float32)
    r = np.full(len(ka), radius, np.float32)
    r[::7] = 0.0                                   # empty windows (what the facade passes for keypoints it skips)
    lo = np.full(len(ka), levels[0], np.int32); hi = np.full(len(ka), levels[1], np.int32)
    okb = np.stack([kb["x"], kb["y"], kb["size"], kb["angle"], kb["response"], kb["octave"].astype(np.float32)], 1)
    F = O.OracleFrame(okb, db, w, h)
    want = [F.features_in_area(qx[i], qy[i], r[i], int(lo[i]), int(hi[i])) if r[i] > 0 else [] for i in range(len(ka))]
    total_want = sum(len(c) for c in want)
    assert total_want > 1000
    kbr = np.ascontiguousarray(kb.astype(KP_DTYPE))
    bounds = np.array([0, 0, w, h], np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)

    def call(cap):
        offs = np.zeros(len(ka) + 1, np.int32); cands = np.zeros(max(cap, 1), np.int32); dist = np.zeros(max(cap, 1), np.int16)
        total = C.c_int(0)
        rc = L.orbm_window_lists(0, p(kbr), len(kbr), p(bounds), p(db), p(da), len(ka), p(qx), p(qy), p(r), p(lo), p(hi), p(offs), p(cands), p(dist),
                                 cap, C.byref(total))
        return rc, total.value, offs, cands, dist

    rc, total, _, _, _ = call(16)                  # too small: the needed room is reported
    assert rc == _lib.ORB_ECAPACITY and total == total_want
    rc, total, offs, cands, dist = call(total_want)
    assert rc == _lib.ORB_OK and total == total_want and offs[-1] == total_want
    for i in range(len(ka)):
        got = cands[offs[i]:offs[i + 1]]
        assert list(got) == list(want[i]), i
        for k, j in enumerate(got[:8]):
            assert dist[offs[i] + k] == O.hamming(da[i], db[j])
