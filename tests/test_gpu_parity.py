"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs. Integer / byte / index results must be bit-exact; angles are compared as bit patterns."""
import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import ORBextractor, ORBmatcher, synth
from multiagent_orb_slam2_b200.extractor import quadtree as gpu_quadtree

pytestmark = pytest.mark.gpu

CONFIGS = {  # name: (w, h, nfeatures, iniTh)   SURVEY.md section 5 / BASELINE.json configs
    "tum": (640, 480, 1000, 20),
    "kitti": (1241, 376, 2000, 20),
    "euroc": (752, 480, 1200, 20),
}


def run_both(w, h, nf, ini, kind, seed, nlevels=8, scale=1.2, mn=7):
    img = synth.image(kind, w, h, seed)
    o = O.OracleExtractor(nf, scale, nlevels, ini, mn)
    ok, od = o(img)
    g = ORBextractor(nf, scale, nlevels, ini, mn)
    gk, gd = g(img)
    return img, o, ok, od, g, gk, gd


def kp_matrix(gk):
    return np.stack([gk["x"], gk["y"], gk["size"], gk["angle"], gk["response"], gk["octave"].astype(np.float32)], 1)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("kind,seed", [("blocks", 0), ("blurnoise", 1), ("noise", 2)])
def test_stages_bit_exact(cfg, kind, seed):
    w, h, nf, ini = CONFIGS[cfg]
    img, o, ok, od, g, gk, gd = run_both(w, h, nf, ini, kind, seed)
    for l in range(8):
        L = o.level(l)
        assert np.array_equal(g.pyramid_level(l), L["img"]), "pyramid level %d" % l
        if L["blur"] is not None:
            assert np.array_equal(g.blurred_level(l), L["blur"]), "blurred level %d" % l
        assert np.array_equal(g.candidates(l), L["cand"]), "FAST candidates level %d" % l
    assert len(gk) == len(ok)
    assert np.array_equal(kp_matrix(gk).view(np.uint32), ok.view(np.uint32)), "keypoints (incl. angle bits)"
    mism = np.flatnonzero((gd != od).any(1))
    assert len(mism) == 0, "descriptor rows differ: %d of %d" % (len(mism), len(od))


def test_tables_match_oracle():
    g = ORBextractor(1200, 1.2, 8, 20, 7, width=752, height=480)
    t = O.OracleExtractor(1200, 1.2, 8).tables()
    assert np.array_equal(g.GetScaleFactors().view(np.uint32), t["scale"].view(np.uint32))
    assert np.array_equal(g.GetInverseScaleFactors().view(np.uint32), t["inv_scale"].view(np.uint32))
    assert np.array_equal(g.GetScaleSigmaSquares().view(np.uint32), t["sigma2"].view(np.uint32))
    assert np.array_equal(g.GetInverseScaleSigmaSquares().view(np.uint32), t["inv_sigma2"].view(np.uint32))
    assert g.features_per_level().tolist() == t["quota"].tolist()
    assert g.GetLevels() == 8


@pytest.mark.parametrize("nf,s,nl,ini,mn", [(500, 1.2, 8, 20, 7), (1500, 1.1, 5, 15, 5), (300, 1.5, 4, 30, 10), (3000, 1.2, 8, 12, 7)])
def test_other_settings(nf, s, nl, ini, mn):
    img, o, ok, od, g, gk, gd = run_both(640, 480, nf, ini, "blocks", 3, nlevels=nl, scale=s, mn=mn)
    assert np.array_equal(kp_matrix(gk).view(np.uint32), ok.view(np.uint32))
    assert np.array_equal(gd, od)


def test_flat_and_empty_images():
    g = ORBextractor(1000, 1.2, 8, 20, 7)
    k, d = g(synth.image("flat", 640, 480, 0))
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = g(np.empty((0, 0), np.uint8))
    assert len(k) == 0


def test_batch_equals_single_frames():
    imgs = np.stack([synth.image(k, 640, 480, s) for k, s in [("blocks", 0), ("blurnoise", 1), ("blocks", 2), ("noise", 3), ("flat", 0)]])
    g = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=8)
    kps, desc, counts = g.extract_batch(imgs)
    o = O.OracleExtractor()
    for i in range(len(imgs)):
        ok, od = o(imgs[i])
        assert counts[i] == len(ok)
        assert np.array_equal(kp_matrix(kps[i, :counts[i]]).view(np.uint32), ok.view(np.uint32)), i
        assert np.array_equal(desc[i, :counts[i]], od), i


@pytest.mark.parametrize("seed", range(4))
def test_quadtree_kernel_on_random_candidate_sets(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(25):
        W, H = [(608, 448), (1209, 344), (720, 448), (147, 102), (314, 73), (1000, 333)][rng.integers(0, 6)]
        n = int(rng.choice([1, 2, 3, 7, 50, 300, 1500, 6000, 30000]))
        N = int(rng.choice([1, 5, 60, 217, 434, 1000]))
        x = rng.integers(0, W - 6, n); y = rng.integers(0, H - 6, n)
        if rng.random() < 0.5:
            x = np.clip(rng.normal(W / 3, 25, n), 0, W - 7).astype(np.int64); y = np.clip(rng.normal(H / 2, 25, n), 0, H - 7).astype(np.int64)
        _, first = np.unique(x * 8192 + y, return_index=True)
        first = np.sort(first)
        x, y = x[first].astype(np.int32), y[first].astype(np.int32)
        sc = rng.integers(7, 255, len(x)).astype(np.int32)
        a = O.quadtree(x, y, sc, 16, 16 + W, 16, 16 + H, N)
        b = gpu_quadtree(x, y, sc, 16, 16 + W, 16, 16 + H, N)
        assert np.array_equal(a, b), (W, H, n, N)


# ---------------------------------------------------------------------------------------------
@pytest.fixture(params=[1, 2, 3], ids=["popc", "tensorcore", "tensorcore-cta-pairs"])
def knn2_backend(request):
    """Run a test once per brute-force implementation (csrc/hamming.cu POPC kernel, csrc/hamming_mma.cu tcgen05 kernel with one
    CTA per query tile, the same with cta_group::2 CTA pairs)."""
    from multiagent_orb_slam2_b200 import _lib
    _lib.check(_lib.lib().orbm_set_knn2_backend(request.param))
    yield request.param
    _lib.check(_lib.lib().orbm_set_knn2_backend(0))


@pytest.mark.parametrize("nA,nB", [(1000, 1000), (2013, 2013), (1, 1), (7, 5000), (3000, 33), (257, 513), (5, 0), (128, 256), (129, 255), (640, 1)])
def test_knn2_bit_exact(nA, nB, knn2_backend):
    B = synth.descriptors(nB, 1)
    A = synth.descriptors(nA, 2, dup_from=B) if nB else synth.descriptors(nA, 2)
    if nB > 10:  # exact duplicates: ties between best and second, first index must win
        B[nB // 2] = B[3]
    m = ORBmatcher(0.75)
    gi, g1, g2 = m.knn2(A, B)
    oi, o1, o2 = O.knn2(A, B)
    assert np.array_equal(gi, oi) and np.array_equal(g1, o1) and np.array_equal(g2, o2)


def test_knn2_large_split_database(knn2_backend):
    B = synth.descriptors(60000, 3)
    A = synth.descriptors_fast(4000, 4, B)
    m = ORBmatcher(0.75)
    gi, g1, g2 = m.knn2(A, B)
    oi, o1, o2 = O.knn2(A, B, threads=8)
    assert np.array_equal(gi, oi) and np.array_equal(g1, o1) and np.array_equal(g2, o2)
    acc = m.accept(gi, g1, g2)
    assert (acc >= 0).sum() > 1000


def test_knn2_split_scan_keeps_first_minimum_across_segments(knn2_backend):
    """Long databases are scanned by several CTAs per query tile (candidate-range split, csrc/hamming_mma.cu) and the partial
    results folded afterwards: exact duplicates of a query's best row that fall into different segments must still resolve to
    the FIRST one, with second best == best (the reference's strict '<', src/ORBmatcher.cc:589-598)."""
    nB, nA = 40000, 300
    B = synth.descriptors(nB, 11)
    src = np.array([5, 9000, 13000, 21000, 33000, 39999])
    A = synth.descriptors_fast(nA, 12, B, 40)
    A[:6] = B[src]                                  # exact hits, one per region of the database
    for k, j in enumerate(src):                     # and later copies of each of them in other segments
        for off in (7001, 15003, 26007):
            if j + off < nB:
                B[j + off] = B[j]
    A[6] = B[39999]
    m = ORBmatcher(0.75)
    gi, g1, g2 = m.knn2(A, B)
    oi, o1, o2 = O.knn2(A, B, threads=8)
    assert np.array_equal(gi, oi) and np.array_equal(g1, o1) and np.array_equal(g2, o2)
    assert g1[0] == 0 and g2[0] == 0 and gi[0] == 5 and gi[1] == 9000


def test_knn2_lists_and_distance_matrix():
    rng = np.random.default_rng(5)
    B = synth.descriptors(1500, 6)
    A = synth.descriptors(800, 7, dup_from=B)
    lens = rng.integers(0, 60, len(A))
    lens[:5] = [0, 1, 2, 33, 64]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cands = rng.integers(0, len(B), offsets[-1]).astype(np.int32)
    m = ORBmatcher()
    gi, g1, g2 = m.knn2_lists(A, B, offsets, cands)
    oi, o1, o2 = O.knn2_lists(A, B, offsets, cands)
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2) and np.array_equal(gi, oi)
    D = m.distance_matrix(A[:100], B[:77])
    ref = np.array([[O.hamming(a, b) for b in B[:77]] for a in A[:100]], np.int16)
    assert np.array_equal(D, ref)


def test_agent_frontend_extract_and_match_consecutive():
    import torch
    from multiagent_orb_slam2_b200.frontend import AgentFrontend
    frames = []
    for s in range(3):
        a, b, _ = synth.shifted_pair("blocks", 640, 480, s)
        frames += [a, b]
    imgs = np.stack(frames)
    fe = AgentFrontend(640, 480, max_batch=8)
    kps, desc, counts, match = fe.process(imgs)
    o = O.OracleExtractor()
    od = []
    for i in range(len(imgs)):
        ok, d = o(imgs[i])
        assert counts[i] == len(ok) and np.array_equal(desc[i, :counts[i]], d)
        od.append(d)
    m = ORBmatcher(0.9)
    n = len(imgs)
    for i in range(n):
        oi, o1, o2 = O.knn2(od[i], od[(i + 1) % n])
        ref = m.accept(oi, o1, o2, th=50, inclusive=True)
        assert np.array_equal(match[i, :counts[i]], ref), i
    assert (match[0, :counts[0]] >= 0).sum() > 200  # shifted copy: plenty of true matches
    # device-resident path gives the same matches
    d_imgs = torch.from_numpy(imgs).cuda()
    fe.process_device(d_imgs)
    torch.cuda.synchronize()
    assert np.array_equal(fe.match[:n].cpu().numpy()[0, :counts[0]], match[0, :counts[0]])


def test_device_frames_with_unaligned_pitch():
    """Caller-owned device frames whose row pitch is not a multiple of 4 (1241 px wide, tightly packed)
    take the byte-wise tile loads of FAST / blur; results must not change."""
    import torch
    img = synth.image("blocks", 1241, 376, 5)
    ok, od = O.OracleExtractor(2000, 1.2, 8, 20, 7)(img)
    g = ORBextractor(2000, 1.2, 8, 20, 7, width=1241, height=376, max_batch=2)
    d = torch.from_numpy(np.stack([img, img])).cuda()
    assert d.stride(1) == 1241
    g.extract_device(d.data_ptr(), d.stride(1), d.stride(0), 2, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    from multiagent_orb_slam2_b200 import _lib
    import ctypes as C
    kps = np.empty((2, g.cap), dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
    desc = np.empty((2, g.cap, 32), np.uint8)
    counts = np.empty(2, np.int32)
    _lib.check(_lib.lib().orbx_download_results(g._h, 2, kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), g.cap,
                                                counts.ctypes.data_as(C.c_void_p), None))
    torch.cuda.synchronize()
    for i in range(2):
        assert counts[i] == len(ok)
        assert np.array_equal(kp_matrix(kps[i, :counts[i]]).view(np.uint32), ok.view(np.uint32))
        assert np.array_equal(desc[i, :counts[i]], od)


def test_knn2_extreme_distances(knn2_backend):
    """Distance 0 and distance 256 (complement: never closer than the initial 256, idx stays -1), all-equal candidates."""
    rng = np.random.default_rng(9)
    A = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    m = ORBmatcher(0.75)
    for B in (~A[:1], np.repeat(~A[:1], 700, 0), np.repeat(A[5:6], 513, 0), np.concatenate([~A[:1], A[:1], A[:1]])):
        B = np.ascontiguousarray(B)
        gi, g1, g2 = m.knn2(A, B)
        oi, o1, o2 = O.knn2(A, B)
        assert np.array_equal(gi, oi) and np.array_equal(g1, o1) and np.array_equal(g2, o2)
    gi, g1, g2 = m.knn2(A[:1], np.ascontiguousarray(~A[:1]))
    assert gi[0] == -1 and g1[0] == 256 and g2[0] == 256


def test_knn2_pairs_all_slice_counts(knn2_backend):
    """orbm_knn2_pairs_device picks 1..8 slices per query depending on the grid size: cover them all."""
    import torch
    import ctypes as C
    from multiagent_orb_slam2_b200 import _lib
    L = _lib.lib()
    for nsets, rows in [(2, 300), (6, 700), (40, 1100), (150, 500), (700, 300)]:
        rng = np.random.default_rng(nsets)
        counts = rng.integers(max(1, rows - 200), rows + 1, nsets).astype(np.int32)
        counts[0] = rows
        base = synth.descriptors(rows, 11)
        sets = np.zeros((nsets, rows, 32), np.uint8)
        for k in range(nsets):
            sets[k, :counts[k]] = synth.descriptors_fast(int(counts[k]), 100 + k, base, 40)
        pairs = np.stack([np.arange(nsets), (np.arange(nsets) + 1) % nsets], 1).astype(np.int32)
        d_sets, d_counts, d_pairs = torch.from_numpy(sets).cuda(), torch.from_numpy(counts).cuda(), torch.from_numpy(pairs).cuda()
        out = [torch.empty((nsets, rows), dtype=torch.int32, device="cuda") for _ in range(3)]
        _lib.check(L.orbm_knn2_pairs_device(C.c_void_p(d_sets.data_ptr()), C.c_void_p(d_counts.data_ptr()), rows, C.c_void_p(d_pairs.data_ptr()),
                                            nsets, *[C.c_void_p(o.data_ptr()) for o in out], None))
        torch.cuda.synchronize()
        gi, g1, g2 = [o.cpu().numpy() for o in out]
        for p in rng.choice(nsets, min(nsets, 6), replace=False):
            a, b = pairs[p]
            oi, o1, o2 = O.knn2(sets[a, :counts[a]], sets[b, :counts[b]])
            n = counts[a]
            assert np.array_equal(gi[p, :n], oi) and np.array_equal(g1[p, :n], o1) and np.array_equal(g2[p, :n], o2), (nsets, p)


@pytest.mark.parametrize("w,h,nf,kind,seed", [(1241, 376, 2000, "blocks", 0), (752, 480, 1200, "blocks", 1), (640, 480, 1000, "blurnoise", 2)])
def test_stereo_matches_bit_exact(w, h, nf, kind, seed):
    """Frame::ComputeStereoMatches: KITTI (fx 718.856, baseline*fx 386.1448) / EuRoC-like settings."""
    from multiagent_orb_slam2_b200.extractor import compute_stereo_matches
    left, right = synth.stereo_pair(kind, w, h, seed)
    mbf, fx = np.float32(386.1448), np.float32(718.856)
    mb = np.float32(mbf / fx)
    oL, oR = O.OracleExtractor(nf, 1.2, 8, 20, 7), O.OracleExtractor(nf, 1.2, 8, 20, 7)
    kL, _ = oL(left)
    oR(right)
    ou, od, okept = O.stereo_match(oL, oR, mbf, mb)
    gL, gR = ORBextractor(nf, 1.2, 8, 20, 7), ORBextractor(nf, 1.2, 8, 20, 7)
    gL(left); gR(right)
    gu, gd, gkept = compute_stereo_matches(gL, gR, mbf, mb)
    n = len(kL)
    assert gkept == okept and okept > 100
    assert np.array_equal(gu[:n].view(np.uint32), ou.view(np.uint32))
    assert np.array_equal(gd[:n].view(np.uint32), od.view(np.uint32))


def test_knn2_tensorcore_on_two_streams_at_once():
    """Two agents' frontends share a GPU on separate streams: the tensor-core matcher keeps its expanded operands per
    stream, so overlapping calls must not disturb each other."""
    import torch
    import ctypes as C
    from multiagent_orb_slam2_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    n = 6000
    data = []
    for k in range(2):
        B = synth.descriptors(n, 70 + k)
        A = synth.descriptors_fast(n, 80 + k, B, 50)
        data.append((torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)))

    def run(backend, streams):
        _lib.check(L.orbm_set_knn2_backend(backend))
        outs = [[torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)] for _ in range(2)]
        torch.cuda.synchronize()
        for rep in range(3):
            for k in range(2):
                A, B = data[k]
                _lib.check(L.orbm_knn2_device(C.c_void_p(A.data_ptr()), n, C.c_void_p(B.data_ptr()), n, *[C.c_void_p(o.data_ptr()) for o in outs[k]],
                                              C.c_void_p(streams[k].cuda_stream)))
        torch.cuda.synchronize()
        _lib.check(L.orbm_set_knn2_backend(0))
        return [[o.cpu().numpy() for o in out] for out in outs]

    s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    want = run(1, [s0, s0])
    got = run(2, [s0, s1])
    for k in range(2):
        for g, w in zip(got[k], want[k]):
            assert np.array_equal(g, w)


def test_tensorcore_tile_dot_products():
    """The data path of the tensor-core matcher on one 128 x 256 tile: bit expansion, SWIZZLE_128B TMA loads, tcgen05.mma
    kind::i8 and tcgen05.ld must reproduce dot = 256 - 2 * Hamming distance for every pair."""
    import ctypes as C
    from multiagent_orb_slam2_b200 import _lib
    rng = np.random.default_rng(12)
    A = rng.integers(0, 256, (128, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (256, 32), dtype=np.uint8)
    B[5] = A[7]; B[200] = ~A[100]
    out = np.zeros((128, 256), np.int32)
    _lib.check(_lib.lib().orbm_debug_mma_dot(0, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    ham = np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(2).astype(np.int32)
    assert np.array_equal(out, 256 - 2 * ham)
    assert out[7, 5] == 256 and out[100, 200] == -256


def test_host_wrappers_from_many_threads():
    """ORBmatcher is called concurrently from the Tracking / LocalMapping / LoopClosing / MapFusion threads: the host-buffer
    entry points keep a workspace and a stream per calling thread."""
    import threading
    B = synth.descriptors(1500, 21)
    sets = [synth.descriptors_fast(900 + 50 * k, 30 + k, B, 50) for k in range(6)]
    want = [O.knn2(a, B) for a in sets]
    errors = []

    def worker(k):
        try:
            m = ORBmatcher(0.75)
            for rep in range(5):
                gi, g1, g2 = m.knn2(sets[k], B)
                D = m.distance_matrix(sets[k][:40], B[:60])
                assert np.array_equal(gi, want[k][0]) and np.array_equal(g1, want[k][1]) and np.array_equal(g2, want[k][2])
                assert D[3, 7] == O.hamming(sets[k][3], B[7])
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_stereo_batch_equals_single_frames_and_oracle():
    """orbm_stereo_match_batch_device (all frames of a batched L / R extraction in one launch pair) against the per-frame
    call and the oracle."""
    import ctypes as C
    import torch
    from multiagent_orb_slam2_b200 import _lib
    from multiagent_orb_slam2_b200.extractor import compute_stereo_matches
    w, h, nf, B = 752, 480, 1200, 5
    pairs = [synth.stereo_pair("blocks", w, h, 20 + s) for s in range(B)]
    left, right = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    mbf, fx = np.float32(386.1448), np.float32(718.856)
    mb = np.float32(mbf / fx)
    gL, gR = ORBextractor(nf, 1.2, 8, 20, 7, w, h, max_batch=B), ORBextractor(nf, 1.2, 8, 20, 7, w, h, max_batch=B)
    kL, _, cL = gL.extract_batch(left)
    gR.extract_batch(right)
    cap = gL.cap
    L = _lib.lib()
    ur = torch.full((B, cap + 3), -7.0, dtype=torch.float32, device="cuda"); dz = torch.full_like(ur, -7.0)
    sad = torch.zeros((B, cap + 3), dtype=torch.int32, device="cuda"); kept = torch.zeros(B, dtype=torch.int32, device="cuda")
    _lib.check(L.orbm_stereo_match_batch_device(gL._h, gR._h, B, float(mbf), float(mb), C.c_void_p(ur.data_ptr()), C.c_void_p(dz.data_ptr()),
                                                C.c_void_p(sad.data_ptr()), C.c_void_p(kept.data_ptr()), cap + 3, None))
    torch.cuda.synchronize()
    for f in range(B):
        su, sd, sk = compute_stereo_matches(gL, gR, mbf, mb, frame=f)
        n = int(cL[f])
        assert int(kept[f]) == sk and sk > 100
        assert np.array_equal(ur[f, :n].cpu().numpy().view(np.uint32), su[:n].view(np.uint32))
        assert np.array_equal(dz[f, :n].cpu().numpy().view(np.uint32), sd[:n].view(np.uint32))
        if f in (0, B - 1):
            oL, oR = O.OracleExtractor(nf, 1.2, 8, 20, 7), O.OracleExtractor(nf, 1.2, 8, 20, 7)
            oL(left[f]); oR(right[f])
            ou, od, ok = O.stereo_match(oL, oR, mbf, mb)
            assert ok == sk and np.array_equal(su[:n].view(np.uint32), ou.view(np.uint32)) and np.array_equal(sd[:n].view(np.uint32), od.view(np.uint32))
