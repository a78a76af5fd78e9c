"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/orb_b200.h declares; without a GPU every compute call fails loudly (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from multiagent_orb_slam2_b200 import _lib


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.exported_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump not available")
    archs = {l.split(".")[-2] for l in out.splitlines() if ".cubin" in l}
    assert archs == {"sm_100a"}, archs


def test_argument_validation_needs_no_gpu():
    L = _lib.lib()
    h = C.c_void_p()
    cfg = _lib.Config(1000, 1.2, 99, 20, 7)
    assert L.orbx_create(C.byref(cfg), 0, 640, 480, 1, C.byref(h)) == _lib.ORB_EINVAL
    assert b"nlevels" in L.orb_last_error()
    cfg = _lib.Config(1000, 1.2, 8, 5, 7)
    assert L.orbx_create(C.byref(cfg), 0, 640, 480, 1, C.byref(h)) == _lib.ORB_EINVAL
    assert L.orbm_knn2_device(None, -1, None, 0, None, None, None, None) == _lib.ORB_EINVAL
    assert L.orbm_set_knn2_backend(4) == _lib.ORB_EINVAL and L.orbm_set_knn2_backend(-1) == _lib.ORB_EINVAL
    assert L.orbm_set_knn2_backend(3) == _lib.ORB_OK and L.orbm_set_knn2_backend(2) == _lib.ORB_OK and L.orbm_set_knn2_backend(0) == _lib.ORB_OK
    assert L.orbm_knn2_mma_device(None, -1, None, 0, None, None, None, None) == _lib.ORB_EINVAL
    db = C.c_void_p()
    assert L.orbdb_create(0, 0, C.byref(db)) == _lib.ORB_EINVAL
    total = C.c_int(7)
    assert L.orbm_window_lists(0, None, 0, None, None, None, 0, None, None, None, None, None, None, None, None, 0, C.byref(total)) == _lib.ORB_OK
    assert total.value == 0                                                                   # no queries: nothing to do, no GPU needed
    assert L.orbm_window_lists(0, None, 5, None, None, None, 3, None, None, None, None, None, None, None, None, 0, C.byref(total)) == _lib.ORB_EINVAL
    assert L.orbm_release_scratch(None) in (_lib.ORB_OK, _lib.ORB_ECUDA)                      # exported; needs a device to do anything


def test_no_cpu_fallback_without_device():
    L = _lib.lib()
    if L.orb_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    cfg = _lib.Config(1000, 1.2, 8, 20, 7)
    assert L.orbx_create(C.byref(cfg), 0, 640, 480, 1, C.byref(h)) == _lib.ORB_ECUDA
    a = np.zeros((4, 32), np.uint8)
    o = np.zeros(4, np.int32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    assert L.orbm_knn2(0, p(a), 4, p(a), 4, p(o), p(o), p(o)) == _lib.ORB_ECUDA
    from multiagent_orb_slam2_b200 import ORBmatcher, OrbError
    with pytest.raises(OrbError):
        ORBmatcher().knn2(a, a)


def test_three_maxima_and_descriptor_distance_host_helpers():
    from multiagent_orb_slam2_b200 import ORBmatcher
    import oracle_lib as O
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
        assert ORBmatcher.DescriptorDistance(a, b) == O.hamming(a, b)
    assert ORBmatcher.ComputeThreeMaxima([0, 30, 3, 2] + [0] * 26) == (1, 2, -1)
    assert ORBmatcher.ComputeThreeMaxima([5, 30, 3, 2] + [0] * 26) == (1, 0, 2)
    assert ORBmatcher.ComputeThreeMaxima([0] * 30) == (-1, -1, -1)


def test_image_bounds_host_entry_equals_oracle():
    """orbm_image_bounds (Frame::ComputeImageBounds, src/Frame.cc:436-464) is host arithmetic: checked here without a GPU
    against the oracle restatement (pinned to cv2.undistortPoints and to the reference's Frame constructor)."""
    import oracle_lib as O
    L = _lib.lib()
    f32 = np.float32
    for K4, dist in [([517.306408, 516.469215, 318.643040, 255.313989], [0.262383, -0.953104, -0.005358, 0.002628, 1.163314]),
                     ([458.654, 457.296, 367.215, 248.375], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]),
                     ([500.0, 500.0, 320.0, 240.0], [0.0, 0.0, 0.0, 0.0])]:
        for w, h in ((640, 480), (752, 480)):
            d = np.array(dist, f32)
            b = np.zeros(4, f32)
            assert L.orbm_image_bounds(w, h, *[float(f32(v)) for v in K4], d.ctypes.data_as(C.c_void_p), len(d), b.ctypes.data_as(C.c_void_p)) == _lib.ORB_OK
            Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], f32)
            assert np.array_equal(b.view(np.uint32), O.image_bounds(w, h, Km, d).view(np.uint32)), (K4, w, h)
    assert L.orbm_image_bounds(640, 480, 500.0, 500.0, 320.0, 240.0, d.ctypes.data_as(C.c_void_p), 3, b.ctypes.data_as(C.c_void_p)) == _lib.ORB_EINVAL
