"""Edge cases of the boundary on a GPU: batch limits, capacity errors, unsupported geometry, handle reuse,
concurrent handles (the reference runs a left and a right extractor in two threads, src/Frame.cc:78-81)."""
import ctypes as C
import threading

import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import ORBextractor, ORBmatcher, OrbError, _lib, synth

pytestmark = pytest.mark.gpu


def test_batch_larger_than_max_batch_is_rejected():
    g = ORBextractor(500, 1.2, 8, 20, 7, max_batch=2)
    with pytest.raises(OrbError):
        g.extract_batch(np.zeros((3, 240, 320), np.uint8))


def test_too_small_image_is_rejected_like_the_reference_would_crash():
    with pytest.raises(OrbError) as e:
        ORBextractor(500, 1.2, 8, 20, 7, width=120, height=100)  # level 7 is 33 px: nCols = 0 in the reference
    assert "smaller than one 30 px cell" in str(e.value)


def test_capacity_error_still_fills_the_buffer():
    img = synth.image("blocks", 640, 480, 0)
    g = ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480)
    L = _lib.lib()
    cap = 100
    kps = np.empty(cap, dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
    desc = np.empty((cap, 32), np.uint8)
    n = C.c_int()
    rc = L.orbx_extract(g._h, C.c_void_p(img.ctypes.data), img.strides[0], kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap, C.byref(n))
    assert rc == _lib.ORB_ECAPACITY and n.value > cap
    ok, od = O.OracleExtractor()(img)
    assert np.array_equal(desc, od[:cap])


def test_handle_reuse_and_image_size_change():
    g = ORBextractor(800, 1.2, 8, 20, 7)
    for (w, h, seed) in [(640, 480, 0), (752, 480, 1), (640, 480, 2)]:
        img = synth.image("blocks", w, h, seed)
        gk, gd = g(img)
        ok, od = O.OracleExtractor(800)(img)
        assert len(gk) == len(ok) and np.array_equal(gd, od)


def test_two_extractors_in_two_threads():
    left, right = synth.stereo_pair("blocks", 752, 480, 4)
    ex = [ORBextractor(1200, 1.2, 8, 20, 7, width=752, height=480) for _ in range(2)]
    out = [None, None]

    def run(i, img):
        for _ in range(5):
            out[i] = ex[i](img)

    th = [threading.Thread(target=run, args=(i, im)) for i, im in enumerate((left, right))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i, im in enumerate((left, right)):
        ok, od = O.OracleExtractor(1200)(im)
        assert np.array_equal(out[i][1], od)


def test_many_features_and_many_levels():
    img = synth.image("noise", 1241, 376, 9)
    for nf, s, nl in [(5000, 1.2, 8), (2000, 1.1, 12)]:
        gk, gd = ORBextractor(nf, s, nl, 20, 7)(img)
        ok, od = O.OracleExtractor(nf, s, nl, 20, 7)(img)
        assert len(gk) == len(ok) and np.array_equal(gd, od)


def test_matcher_empty_and_degenerate_sets():
    m = ORBmatcher(0.8)
    A = synth.descriptors(10, 1)
    idx, d1, d2 = m.knn2(A, np.empty((0, 32), np.uint8))
    assert (idx == -1).all() and (d1 == 256).all() and (d2 == 256).all()
    idx, d1, d2 = m.knn2(A, A[:1])
    assert (idx == 0).all() and (d2 == 256).all() and d1[0] == 0
    comp = (~A[:1]).copy()  # complement: distance 256 is not < 256 => no candidate, as in the reference
    idx, d1, d2 = m.knn2(A[:1], comp)
    assert idx[0] == -1 and d1[0] == 256


@pytest.mark.parametrize("w,h,nf", [(1920, 1080, 3000), (480, 640, 800), (3000, 700, 2500)])
def test_other_image_geometries(w, h, nf):
    """Full HD, portrait (aspect < 1: a single quadtree root) and a wide panorama (4 roots)."""
    img = synth.image("blocks", w, h, 77)
    gk, gd = ORBextractor(nf, 1.2, 8, 20, 7)(img)
    ok, od = O.OracleExtractor(nf, 1.2, 8, 20, 7)(img)
    got = np.stack([gk["x"], gk["y"], gk["size"], gk["angle"], gk["response"], gk["octave"].astype(np.float32)], 1)
    assert len(gk) == len(ok) and np.array_equal(got.view(np.uint32), ok.view(np.uint32)) and np.array_equal(gd, od)


def _plateau_image(w, h, seed):
    """2x2-replicated noise: neighbouring pixels with EQUAL FAST scores, so that cells exist whose corners at
    iniThFAST all lose the strict 3x3 maximum test (the cell is then empty and is redone at minThFAST)."""
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
    return np.ascontiguousarray(np.kron(small, np.ones((2, 2), np.uint8))[:h, :w])


@pytest.mark.parametrize("ini,mn", [(20, 7), (12, 12), (60, 5), (120, 1), (254, 1)])
@pytest.mark.parametrize("kind", ["blocks", "blurnoise", "plateau", "flat"])
def test_fast_thresholds_and_empty_cell_retry(ini, mn, kind):
    """The FAST kernel runs every cell at iniThFAST and repeats it at minThFAST only when it comes back empty
    (src/ORBextractor.cc:809-816): candidates of every level, keypoints and descriptors against the oracle for
    thresholds that make the retry rare, frequent, universal, or a no-op (ini == min)."""
    w, h = 640, 480
    img = _plateau_image(w, h, 5) if kind == "plateau" else synth.image(kind, w, h, 5)
    o = O.OracleExtractor(1000, 1.2, 8, ini, mn)
    ok, od = o(img)
    g = ORBextractor(1000, 1.2, 8, ini, mn)
    gk, gd = g(img)
    for l in range(8):
        assert np.array_equal(g.candidates(l), o.level(l)["cand"]), "FAST candidates level %d" % l
    got = np.stack([gk["x"], gk["y"], gk["size"], gk["angle"], gk["response"], gk["octave"].astype(np.float32)], 1)
    assert len(gk) == len(ok)
    if len(ok):
        assert np.array_equal(got.view(np.uint32), ok.view(np.uint32)) and np.array_equal(gd, od)


def test_three_frontends_in_flight_share_one_compute_stream():
    """The end-to-end pipeline of bench.py: three agents' buffers in flight on three streams, their kernels
    serialised on the GPU's shared compute stream (frontend.py). Every step's results must equal those of
    the same frames processed alone."""
    import torch
    from multiagent_orb_slam2_b200.frontend import AgentFrontend
    sets = []
    for k in range(3):
        frames = []
        for s in range(2):
            a, b, _ = synth.shifted_pair("blocks", 640, 480, 10 * k + s)
            frames += [a, b]
        sets.append(torch.from_numpy(np.stack(frames)).pin_memory())
    fes = [AgentFrontend(640, 480, max_batch=4) for _ in range(3)]
    ref = []
    for k in range(3):
        kps, desc, counts, match = fes[k].process(sets[k])
        ref.append((kps.copy(), desc.copy(), counts.copy(), match.copy()))
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = [f.pinned_outputs() for f in fes]
    for rep in range(4):
        for k in range(3):
            with torch.cuda.stream(streams[k]):
                fes[k].process_async(sets[(k + rep) % 3], outs[k])
        for k in range(3):
            streams[k].synchronize()
            rk, rd, rc, rm = ref[(k + rep) % 3]
            assert np.array_equal(outs[k]["counts"].numpy()[:4], rc)
            for i in range(4):
                c = rc[i]
                assert np.array_equal(outs[k]["desc"].numpy()[i, :c], rd[i, :c])
                assert np.array_equal(outs[k]["match"].numpy()[i, :c], rm[i, :c])


@pytest.mark.parametrize("w,h,nl", [(91, 91, 1), (95, 150, 2), (150, 93, 2)])
def test_single_large_cell_levels(w, h, nl):
    """Levels only 59..63 px wide inside the border are ONE cell of up to 63 px (nCols = int(width / 30) = 1 needs
    width < 60; 2 cells of 30..44 px otherwise): the largest FAST tiles there are, more than 48 KB of shared memory
    per block and two mask words per row."""
    img = synth.image("blocks", w, h, 3)
    o = O.OracleExtractor(300, 1.2, nl, 20, 7)
    ok, od = o(img)
    g = ORBextractor(300, 1.2, nl, 20, 7)
    gk, gd = g(img)
    for l in range(nl):
        assert np.array_equal(g.candidates(l), o.level(l)["cand"]), "FAST candidates level %d" % l
    got = np.stack([gk["x"], gk["y"], gk["size"], gk["angle"], gk["response"], gk["octave"].astype(np.float32)], 1)
    assert len(gk) == len(ok)
    if len(ok):
        assert np.array_equal(got.view(np.uint32), ok.view(np.uint32)) and np.array_equal(gd, od)


def test_alternating_device_input_buffers_keep_their_graphs():
    """A caller that double-buffers its device frames alternates two FrameSets on one handle: both get their own captured
    graph (small LRU), results stay those of the frame actually given - also for a third and fifth buffer (LRU eviction)."""
    import torch
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=2)
    imgs = [synth.image("blocks", 640, 480, 70 + i) for i in range(5)]
    want = [O.OracleExtractor()(im) for im in imgs]
    bufs = [torch.from_numpy(im).cuda() for im in imgs]
    st = torch.cuda.Stream()
    import ctypes as C
    from multiagent_orb_slam2_b200 import _lib
    from multiagent_orb_slam2_b200.extractor import KP_DTYPE
    L = _lib.lib()
    order = [0, 1, 0, 1, 0, 1, 2, 0, 3, 4, 1, 2, 2, 2, 0, 1]
    for i in order:
        with torch.cuda.stream(st):
            ex.extract_device(bufs[i].data_ptr(), 640, 640 * 480, 1, st.cuda_stream)
        st.synchronize()
        kps = np.empty(ex.cap, KP_DTYPE); desc = np.empty((ex.cap, 32), np.uint8); cnt = np.zeros(1, np.int32)
        _lib.check(L.orbx_download_results(ex._h, 1, kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), ex.cap,
                                           cnt.ctypes.data_as(C.c_void_p), C.c_void_p(st.cuda_stream)))
        st.synchronize()
        n = int(cnt[0])
        ok, od = want[i]
        assert n == len(ok) and np.array_equal(desc[:n], od), (i, n, len(ok))
        assert np.array_equal(kps["x"][:n].view(np.uint32), ok[:, 0].view(np.uint32))


def test_matcher_scratch_grows_is_released_and_survives_thread_exit():
    """The tensor-core matcher keeps its expanded operands per (device, stream). Growing it between calls queued on the same
    stream (no synchronisation in between) must not disturb the earlier call; orbm_release_scratch hands the memory back
    stream-ordered and the next call allocates again; a host-call thread that exits takes its stream's scratch with it."""
    import torch
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(dev)
    sizes = [1500, 4000, 2500, 9000]
    sets = []
    for k, n in enumerate(sizes):
        B = synth.descriptors(n, 300 + k)
        A = synth.descriptors_fast(n, 400 + k, B, 45)
        sets.append((A, B, torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)))
    want = []
    _lib.check(L.orbm_set_knn2_backend(1))
    for A, B, dA, dB in sets:
        o = [torch.empty(len(A), dtype=torch.int32, device=dev) for _ in range(3)]
        _lib.check(L.orbm_knn2_device(C.c_void_p(dA.data_ptr()), len(A), C.c_void_p(dB.data_ptr()), len(B), *[C.c_void_p(x.data_ptr()) for x in o], None))
        torch.cuda.synchronize()
        want.append([x.cpu().numpy() for x in o])
    _lib.check(L.orbm_set_knn2_backend(2))
    try:
        for rnd in range(2):
            outs = []
            for A, B, dA, dB in sets:      # queued back to back: every second call outgrows the scratch of the one before
                o = [torch.empty(len(A), dtype=torch.int32, device=dev) for _ in range(3)]
                _lib.check(L.orbm_knn2_device(C.c_void_p(dA.data_ptr()), len(A), C.c_void_p(dB.data_ptr()), len(B), *[C.c_void_p(x.data_ptr()) for x in o],
                                              C.c_void_p(st.cuda_stream)))
                outs.append(o)
            _lib.check(L.orbm_release_scratch(C.c_void_p(st.cuda_stream)))   # ordered after the queued calls
            st.synchronize()
            for o, w in zip(outs, want):
                for g, ww in zip(o, w):
                    assert np.array_equal(g.cpu().numpy(), ww)
        _lib.check(L.orbm_release_scratch(C.c_void_p(st.cuda_stream)))       # nothing left: still fine
        free_before = torch.cuda.mem_get_info(0)[0]
        errs = []

        def host_call(k):
            try:
                _lib.check(L.orbm_set_knn2_backend(2))
                A, B = sets[k][0], sets[k][1]
                idx, b1, b2 = ORBmatcher().knn2(A, B)
                if not (np.array_equal(idx, want[k][0]) and np.array_equal(b1, want[k][1]) and np.array_equal(b2, want[k][2])):
                    errs.append(k)
            except Exception as e:  # noqa: BLE001
                errs.append(repr(e))
        for rnd in range(3):
            ts = [threading.Thread(target=host_call, args=(k,)) for k in range(4)]
            [t.start() for t in ts]
            [t.join() for t in ts]
        assert not errs, errs
        torch.cuda.synchronize()
        free_after = torch.cuda.mem_get_info(0)[0]
        # 12 exited threads, each with up to 2 x 9000 x 256 B of scratch plus a workspace: none of it may stay behind
        assert free_before - free_after < (8 << 20), (free_before, free_after)
    finally:
        _lib.check(L.orbm_set_knn2_backend(0))


def test_small_batches_take_the_level_parallel_path_and_agree_with_large_ones():
    """Batches of 1 and 2 frames run FAST + quadtree of every pyramid level as its own branch next to the resize chain
    (programmatic dependent launches, captured into a graph from the second identical call on); larger batches run the plain
    sequence. The same frames must give the same keypoints and descriptors whichever way they are batched, call after call."""
    imgs = np.stack([synth.image(kind, 640, 480, 90 + i) for i, kind in enumerate(["blocks", "noise", "blocks", "blurnoise", "blocks", "flat"])])
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=6)
    want_k, want_d, want_c = ex.extract_batch(imgs)            # plain sequence, 6 frames
    for i, im in enumerate(imgs[:3]):                           # the reference itself for three of them
        ok, od = O.OracleExtractor()(im)
        assert want_c[i] == len(ok) and np.array_equal(want_d[i, :want_c[i]], od)
    for rep in range(3):                                        # plain launches, capture, graph replay
        for i in range(6):
            k, d, c = ex.extract_batch(imgs[i:i + 1])
            assert c[0] == want_c[i] and np.array_equal(d[0, :c[0]], want_d[i, :c[0]]) and np.array_equal(k[0, :c[0]], want_k[i, :c[0]]), (rep, i)
        for i in range(0, 6, 2):
            k, d, c = ex.extract_batch(imgs[i:i + 2])
            for j in range(2):
                n = want_c[i + j]
                assert c[j] == n and np.array_equal(d[j, :n], want_d[i + j, :n]) and np.array_equal(k[j, :n], want_k[i + j, :n]), (rep, i, j)
        k, d, c = ex.extract_batch(imgs[:3])
        assert np.array_equal(c, want_c[:3]) and all(np.array_equal(d[j, :c[j]], want_d[j, :c[j]]) for j in range(3))
