"""Frame::isInFrustum / the Fuse prologue on the device (orbm_project_points_device) against the scalar
restatement (oracle_lib.project_points, pinned to cv2 arithmetic in tests/test_projection_oracle.py), and the whole
projection -> window -> Hamming -> ordered replay chain of Tracking::SearchLocalPoints against the oracle."""
import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import ORBextractor, ORBmatcher, synth
from multiagent_orb_slam2_b200 import projection as PJ

pytestmark = pytest.mark.gpu
W, H = 640, 480
f32 = np.float32


def _rotation(rng, max_angle):
    a = rng.uniform(-max_angle, max_angle, 3)
    cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return (Rz @ Ry @ Rx).astype(f32)


def _camera(rng, scale):
    R = _rotation(rng, 0.2)
    t = rng.normal(0, 0.5, 3).astype(f32)
    Ow = (-(R.T.astype(np.float64) @ t.astype(np.float64))).astype(f32)  # mOw = -mRcw.t()*mtcw (double accumulation in cv::gemm)
    cam = dict(Rcw=R, tcw=t, Ow=Ow, fx=f32(517.3), fy=f32(516.5), cx=f32(318.6), cy=f32(255.3), bf=f32(40.0), min_x=f32(0), max_x=f32(W),
               min_y=f32(0), max_y=f32(H), log_scale_factor=np.log(scale[1], dtype=f32), scale=scale)
    c = PJ.Camera.make(R, t, Ow, cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"], (0, W, 0, H), scale)
    assert f32(c.log_scale_factor) == cam["log_scale_factor"]
    return cam, c


def _map_points(rng, cam, kps, n_extra=300):
    """Map points that project near the keypoints (back-projected at random depth, then perturbed) plus points anywhere
    (behind the camera, outside the image, out of scale range, grazing normals) so that every gate fires."""
    R, t = cam["Rcw"].astype(np.float64), cam["tcw"].astype(np.float64)
    n = len(kps)
    z = rng.uniform(0.5, 8.0, n)
    xc = (kps["x"] + rng.normal(0, 1.0, n) - cam["cx"]) / cam["fx"] * z
    yc = (kps["y"] + rng.normal(0, 1.0, n) - cam["cy"]) / cam["fy"] * z
    Pc = np.stack([xc, yc, z], 1)
    Pc = np.concatenate([Pc, rng.normal(0, 4, (n_extra, 3))])
    Pw = (Pc - t) @ R  # R^T (Pc - t)
    Ow = cam["Ow"].astype(np.float64)
    po = Pw - Ow  # MapPoint::mNormalVector is the mean unit vector camera -> point (src/MapPoint.cc:338-374)
    dist = np.linalg.norm(po, axis=1)
    nrm = po / dist[:, None]
    nrm += rng.normal(0, 0.6, nrm.shape)  # spread the viewing angle across the 60 degree gate
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    octave = np.concatenate([kps["octave"], rng.integers(0, 8, n_extra)])
    max_d = dist * cam["scale"][octave] * rng.uniform(0.6, 1.5, len(dist))
    min_d = max_d / cam["scale"][-1]
    return Pw.astype(f32), nrm.astype(f32), max_d.astype(f32), min_d.astype(f32)


@pytest.mark.parametrize("seed,mode,th", [(0, 0, 1.0), (1, 0, 3.0), (2, 1, 3.0), (3, 1, 4.0)])
def test_project_points_matches_oracle(seed, mode, th):
    rng = np.random.default_rng(seed)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    k, d = ex(synth.image("blocks", W, H, seed + 70))
    scale = ex.GetScaleFactors()
    cam, c = _camera(rng, scale)
    pos, nrm, max_d, min_d = _map_points(rng, cam, k)
    desc = synth.descriptors(len(pos), seed)
    want = O.project_points(cam, pos, nrm, max_d, min_d, mode, 0.5, th)
    got = PJ.project(0, c, PJ.MapPointArrays(0, pos, nrm, max_d, min_d, desc), mode, 0.5, th)
    got = {k2: v.cpu().numpy() for k2, v in got.items()}
    alive = want["alive"].astype(bool)
    assert 0.2 < alive.mean() < 0.95  # the gates reject some and keep some
    assert np.array_equal(got["alive"], want["alive"])
    for key in ("u", "v", "ur", "view_cos"):
        assert np.array_equal(got[key].view(np.uint32), want[key].view(np.uint32)), key  # bit patterns, NaN included
    assert np.array_equal(got["level"][alive], want["level"][alive])
    assert np.array_equal(got["radius"].view(np.uint32), want["radius"].view(np.uint32))
    assert np.array_equal(got["min_level"][alive], want["level"][alive] - 1) and np.array_equal(got["max_level"], got["level"])
    assert len(set(want["level"][alive])) >= 6


@pytest.mark.parametrize("seed,th", [(0, 1.0), (1, 3.0)])
def test_search_local_points_device_chain_matches_oracle(seed, th):
    from multiagent_orb_slam2_b200.device_grid import DeviceFrameGrid
    rng = np.random.default_rng(seed + 5)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    k, d = ex(synth.image("blocks", W, H, seed + 80))
    scale = ex.GetScaleFactors()
    cam, c = _camera(rng, scale)
    pos, nrm, max_d, min_d = _map_points(rng, cam, k)
    n_mp = len(pos)
    # descriptors of the map points: noisy copies of the keypoint they came from, random for the extra points
    bits = np.unpackbits(np.concatenate([d, synth.descriptors(n_mp - len(d), seed)]), axis=1)
    for i in range(n_mp):
        bits[i, rng.choice(256, rng.integers(0, 60), replace=False)] ^= 1
    mp_desc = np.packbits(bits, axis=1)
    bad = rng.random(n_mp) < 0.05
    uright = np.where(rng.random(len(k)) < 0.5, k["x"] - rng.uniform(2, 40, len(k)), -1).astype(f32)
    occupied = rng.random(len(k)) < 0.1
    # oracle: scalar projection, then the scalar search on its outputs
    pr = O.project_points(cam, pos, nrm, max_d, min_d, 0, 0.5, th)
    ok = np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"].astype(f32)], 1)
    F = O.OracleFrame(ok, d, W, H)
    F.scale, F.uright, F.occupied = scale, uright, occupied
    mp = dict(in_view=pr["alive"].astype(bool), bad=bad, level=pr["level"], view_cos=pr["view_cos"], proj_x=pr["u"], proj_y=pr["v"],
              proj_xr=pr["ur"], desc=mp_desc)
    onm, oas = O.search_by_projection_frame_mappoints(F, mp, th, 0.8)
    gnm, gas, _ = ORBmatcher(0.8).SearchLocalPoints_device(DeviceFrameGrid(ex), c, PJ.MapPointArrays(0, pos, nrm, max_d, min_d, mp_desc),
                                                          bad, k["octave"], uright, occupied, th)
    assert gnm == onm and np.array_equal(gas, oas)
    assert onm > 100


def test_stereo_from_rgbd_matches_restatement():
    """Frame::ComputeStereoFromRGBD (src/Frame.cc:643-664): depth sampled at the truncated keypoint position."""
    from multiagent_orb_slam2_b200.extractor import compute_stereo_from_rgbd
    rng = np.random.default_rng(4)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    k, _ = ex(synth.image("blocks", W, H, 90))
    depth = rng.uniform(0.3, 8.0, (H, W)).astype(f32)
    depth[rng.random((H, W)) < 0.2] = 0.0     # holes of the sensor
    depth[rng.random((H, W)) < 0.02] = -1.0
    mbf = f32(40.0)
    ur, dz = compute_stereo_from_rgbd(ex, depth, mbf)
    n = len(k)
    want_d = np.full(len(ur), -1, f32); want_u = np.full(len(ur), -1, f32)
    for i in range(n):  # scalar restatement of the reference loop
        d = depth[int(k["y"][i]), int(k["x"][i])]
        if d > 0:
            want_d[i] = d
            want_u[i] = f32(k["x"][i] - f32(mbf / d))
    assert np.array_equal(dz.view(np.uint32), want_d.view(np.uint32)) and np.array_equal(ur.view(np.uint32), want_u.view(np.uint32))
    assert 0.6 < (want_d[:n] > 0).mean() < 0.9


def test_undistort_keypoints_grid_and_rgbd_with_tum1_distortion():
    """SURVEY section 8f-2 with a distorted camera (TUM1.yaml): Frame::UndistortKeyPoints / ComputeImageBounds on the device ==
    oracle restatement (itself pinned to cv2 and to the reference's Frame constructor, tests/test_oracle_vs_reference_matcher.py);
    the device grid then buckets mvKeysUn inside the undistorted bounds, and ComputeStereoFromRGBD uses the undistorted x."""
    from multiagent_orb_slam2_b200.device_grid import DeviceFrameGrid
    from multiagent_orb_slam2_b200.extractor import compute_stereo_from_rgbd
    K4 = [517.306408, 516.469215, 318.643040, 255.313989]
    dist = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], f32)
    Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], f32)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    k, d = ex(synth.image("blocks", W, H, 91))
    n = len(k)
    grid = DeviceFrameGrid(ex, K=K4, dist_coef=dist)
    un = grid.undistorted_keypoints(n)
    want = O.undistort_points(np.stack([k["x"], k["y"]], 1), Km, dist)
    assert np.array_equal(np.stack([un["x"], un["y"]], 1).view(np.uint32), want.view(np.uint32))
    for fld in ("size", "angle", "response", "octave"):
        assert np.array_equal(un[fld], k[fld])
    assert np.abs(want - np.stack([k["x"], k["y"]], 1)).max() > 1.0
    bounds = O.image_bounds(W, H, Km, dist)
    assert np.array_equal(np.array(grid.bounds, f32).view(np.uint32), bounds.view(np.uint32))
    # window search over the undistorted grid == oracle frame with mvKeysUn and the undistorted bounds
    okun = np.stack([un["x"], un["y"], k["size"], k["angle"], k["response"], k["octave"].astype(f32)], 1)
    F = O.OracleFrame(okun, d, W, H, bounds=bounds)
    rng = np.random.default_rng(2)
    nq = 300
    qx, qy = rng.uniform(bounds[0] - 5, bounds[1] + 5, nq).astype(f32), rng.uniform(bounds[2] - 5, bounds[3] + 5, nq).astype(f32)
    r = rng.choice([5.0, 20.0, 60.0], nq).astype(f32)
    qd = d[rng.integers(0, n, nq)]
    gi, g1, g2, _, _ = grid.window_knn2(qd, qx, qy, r, 0, 4)
    hits = 0
    for i in range(nq):
        cand = F.features_in_area(qx[i], qy[i], r[i], 0, 4)
        ds = [O.hamming(qd[i], d[j]) for j in cand]
        wantq = (cand[int(np.argmin(ds))], min(ds)) if cand else (-1, 256)
        assert (gi[i], g1[i]) == wantq, i
        hits += len(cand) > 0
    assert hits > 100
    # RGB-D: depth at the raw position, right coordinate from the undistorted x
    depth = rng.uniform(0.3, 8.0, (H, W)).astype(f32)
    depth[rng.random((H, W)) < 0.2] = 0.0
    ur, dz = compute_stereo_from_rgbd(ex, depth, 40.0, d_kps_un=grid.d_kps_un)
    kp6 = np.stack([k["x"], k["y"]], 1)
    wu, wd = O.stereo_from_rgbd(kp6, want, depth, 40.0)
    assert np.array_equal(ur[:n].view(np.uint32), wu.view(np.uint32)) and np.array_equal(dz[:n].view(np.uint32), wd.view(np.uint32))
    # zero distortion: mvKeysUn = mvKeys, bounds = image
    g0 = DeviceFrameGrid(ex, K=K4, dist_coef=[0, 0, 0, 0])
    assert g0.d_kps_un == g0.d_kps and g0.bounds == (0.0, float(W), 0.0, float(H))
