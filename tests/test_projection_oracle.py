"""CPU: the cv::Mat arithmetic the projection oracle restates (oracle_lib._mat3_vec / _norm3) is pinned to the only
OpenCV in the image, cv2 4.13: Rcw*P+tcw through cv2.gemm (small-matrix float path) and cv::norm through cv2.norm."""
import numpy as np
import pytest

import oracle_lib as O

cv2 = pytest.importorskip("cv2")


def test_mat3_vec_and_norm_equal_cv2():
    rng = np.random.default_rng(0)
    for _ in range(3000):
        R = rng.normal(0, 1, (3, 3)).astype(np.float32)
        x = rng.normal(0, 5, (3, 1)).astype(np.float32)
        t = rng.normal(0, 2, (3, 1)).astype(np.float32)
        want = cv2.gemm(R, x, 1.0, None, 0.0) + t  # cv::Mat operator* then operator+
        got = np.array(O._mat3_vec(R, x[:, 0], t[:, 0]), np.float32)
        assert np.array_equal(got, want[:, 0])
        assert O._norm3(x[:, 0]) == np.float32(cv2.norm(x))


def test_predict_scale_levels():
    lsf = np.log(np.float32(1.2), dtype=np.float32)
    assert O.predict_scale(10.0, 10.0, lsf, 8) == 0
    assert O.predict_scale(10.0, 100.0, lsf, 8) == 0       # ratio < 1: negative, clamped
    assert O.predict_scale(10.0, 0.01, lsf, 8) == 7        # clamped to the top level
    assert O.predict_scale(12.5, 10.0, lsf, 8) == 2        # log(1.25)/log(1.2) = 1.22 -> ceil 2
