"""Pins oracle/cvprim.h (the restated OpenCV primitives) bit-exactly to cv2 4.13.0, the only
OpenCV in the image. CPU only."""
import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import synth

cv2 = pytest.importorskip("cv2")

SHAPES = [(640, 480), (1241, 376), (752, 480), (97, 61)]


def level_sizes(w, h, n=8, s=1.2):
    o = O.OracleExtractor(1000, s, n)
    inv = o.tables()["inv_scale"]
    return [(int(np.rint(np.float32(w) * inv[l])), int(np.rint(np.float32(h) * inv[l]))) for l in range(n)]


@pytest.mark.parametrize("w,h", SHAPES)
@pytest.mark.parametrize("kind", ["noise", "blurnoise"])
def test_resize_chain_matches_cv2(w, h, kind):
    img = synth.image(kind, w, h, 1)
    cur_o, cur_c = img, img
    for (lw, lh) in level_sizes(w, h)[1:]:
        nxt_c = cv2.resize(cur_c, (lw, lh), interpolation=cv2.INTER_LINEAR)
        nxt_o = O.resize(cur_o, lw, lh)
        assert np.array_equal(nxt_o, nxt_c), (lw, lh)
        cur_o, cur_c = nxt_o, nxt_c


@pytest.mark.parametrize("w,h", SHAPES + [(45, 39), (8, 9)])
@pytest.mark.parametrize("kind", ["noise", "blurnoise", "blocks"])
def test_gaussian_matches_cv2(w, h, kind):
    img = synth.image(kind, w, h, 2)
    ref = cv2.GaussianBlur(img.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(O.gaussian(img), ref)


def test_border_matches_cv2():
    img = synth.image("noise", 77, 55, 3)
    ref = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    assert np.array_equal(O.border(img, 19), ref)


@pytest.mark.parametrize("th", [20, 7, 12])
def test_fast_cells_match_cv2(th):
    rng = np.random.default_rng(5)
    total = 0
    for kind in ("blocks", "blurnoise", "noise"):
        img = synth.image(kind, 320, 240, 11)
        for _ in range(40):
            cw, ch = rng.integers(7, 66, 2)
            x0, y0 = rng.integers(0, 320 - cw), rng.integers(0, 240 - ch)
            cell = img[y0:y0 + ch, x0:x0 + cw]
            det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                                 type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kps = det.detect(np.ascontiguousarray(cell))
            ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
            got = O.fast(cell, th)
            assert np.array_equal(got, ref), (kind, x0, y0, cw, ch)
            total += len(ref)
    assert total > 500


def test_fast_tiny_cells_are_empty():
    img = synth.image("noise", 64, 64, 0)
    for cw, ch in [(6, 30), (30, 6), (3, 3), (7, 6)]:
        assert len(O.fast(img[:ch, :cw], 7)) == 0


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(0)
    m = rng.integers(-(1 << 21), 1 << 21, (20000, 2))
    m[:8] = [[0, 0], [0, -5], [5, 0], [-5, 0], [0, 5], [7, 7], [-7, 7], [3, -3]]
    for y, x in m:
        assert O.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x))), (y, x)


def test_round_half_even_matches_cvround():
    for v in [0.5, 1.5, 2.5, -0.5, -1.5, 3.4999, 17.5, -17.5]:
        assert O.lib().orc_round(v) == int(np.rint(np.float32(v)))


# ---- the float cv::Mat arithmetic of the stand-in the reference's matcher / Frame code is compiled against -------------
def test_shim_matrix_expressions_match_cv2():
    """oracle/cvprim_mat.h (used by oracle/shim's MatExpr model): gemm in the shapes the reference uses - Rcw*P+tcw,
    -R.t()*t, s*R products, 4x4 poses - cv::norm, and cv::undistortPoints with the TUM1 / 4-coefficient distortions."""
    import ref_slam as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_slam.so not built")
    rng = np.random.default_rng(0)
    f = np.float32
    for _ in range(2000):
        A = rng.normal(0, 1, (3, 3)).astype(f); x = rng.normal(0, 5, (3, 1)).astype(f); t = rng.normal(0, 2, (3, 1)).astype(f)
        B = rng.normal(0, 1, (3, 3)).astype(f)
        assert np.array_equal(cv2.gemm(A, x, 1.0, t, 1.0), R.gemm32f(A, x, 1.0, t, 1.0))
        assert np.array_equal(cv2.gemm(A, t, -1.0, None, 0.0, flags=cv2.GEMM_1_T), R.gemm32f(A, t, -1.0, None, 0.0, 1))
        assert np.array_equal(cv2.gemm(A, B, 1.0, None, 0.0), R.gemm32f(A, B))
        assert np.array_equal(cv2.gemm(A, x, -1.0, None, 0.0), R.gemm32f(A, x, -1.0))
        T = rng.normal(0, 1, (4, 4)).astype(f); c = rng.normal(0, 1, (4, 1)).astype(f)
        assert np.array_equal(cv2.gemm(T, c, 1.0, None, 0.0), R.gemm32f(T, c))
        assert np.array_equal(cv2.gemm(T, T.copy(), 1.0, None, 0.0), R.gemm32f(T, T))
        assert R.norm_l2(x) == cv2.norm(x)
    K = np.array([[517.306408, 0, 318.643040], [0, 516.469215, 255.313989], [0, 0, 1]], f)
    pts = np.stack([rng.uniform(-5, 645, 4000), rng.uniform(-5, 485, 4000)], 1).astype(f)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]
    for d in ([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], [0.1, 0, 0, 0]):
        d = np.array(d, f)
        want = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2)
        assert np.array_equal(R.undistort_points(pts, K, d).view(np.uint32), want.view(np.uint32))
        assert np.array_equal(O.undistort_points(pts[:300], K, d).view(np.uint32), want[:300].view(np.uint32))
