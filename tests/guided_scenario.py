"""Shared by tests/test_gpu_cpp_guided.py (C++ facade == oracle, on the GPU) and tests/test_oracle_vs_reference_matcher.py
(oracle == the reference's own ORBmatcher, on the CPU): one synthetic two-view scene - keypoints of a shifted image pair,
camera poses consistent with the shift, map points, associations, feature vectors - and the results of the nine guided
searches on it computed with the scalar oracle (oracle/oracle_lib.py)."""
import numpy as np

import oracle_lib as O
from multiagent_orb_slam2_b200 import synth

W, H = 640, 480
f32 = np.float32
TH = np.array([3.0, 15.0, 10.0, 10.0, 3.0, 4.0, 7.5, 0.0], f32)


def extract_pair_cpu(seed, kind="blocks"):
    """The two views extracted with the CPU oracle extractor (bit-identical to the product's and the reference's output)."""
    a, b, shift = synth.shifted_pair(kind, W, H, seed)
    ex = O.OracleExtractor(1000, 1.2, 8, 20, 7)
    out = []
    for img in (a, b):
        k, d = ex(img)
        out.append((k.astype(f32).copy(), d.copy()))
    return out, shift, ex.tables()["scale"].astype(f32)


def make_scenario(seed, views, shift, scale):
    rng = np.random.default_rng(seed)
    (k0, d0), (k1, d1) = views
    dx, dy = shift
    n0, n1 = len(k0), len(k1)
    cam = dict(fx=f32(517.3), fy=f32(516.5), cx=f32(318.6), cy=f32(255.3), bf=f32(40.0), mb=f32(0.08), min_x=f32(0), max_x=f32(W), min_y=f32(0),
               max_y=f32(H), log_scale_factor=np.log(scale[1], dtype=f32), scale=scale)
    Z = 4.0
    # view 0 pose: small rotation; view 1 = view 0 moved so that the image content shifts by (-dx, -dy), plus a little forward motion
    a = rng.uniform(-0.1, 0.1, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    R0 = Rz @ Ry @ Rx
    t0 = rng.normal(0, 0.3, 3)
    trel = np.array([-dx * Z / float(cam["fx"]), -dy * Z / float(cam["fy"]), -0.12 if seed % 2 else 0.12])
    T0 = np.eye(4); T0[:3, :3] = R0; T0[:3, 3] = t0
    T1 = np.eye(4); T1[:3, 3] = trel
    T1 = T1 @ T0
    T0, T1 = T0.astype(f32), T1.astype(f32)
    # map points: back-projection of 85 % of the view 0 keypoints at depth Z, plus extras anywhere in front of the cameras
    own = np.flatnonzero(rng.random(n0) < 0.85)
    Xc = np.stack([(k0[own, 0] - cam["cx"]) / cam["fx"] * Z, (k0[own, 1] - cam["cy"]) / cam["fy"] * Z, np.full(len(own), Z)], 1).astype(np.float64)
    n_extra = 250
    Xe = np.stack([rng.uniform(-3, 3, n_extra), rng.uniform(-2.5, 2.5, n_extra), rng.uniform(-1, 9, n_extra)], 1)
    Xc = np.concatenate([Xc, Xe])
    pos = ((Xc - T0[:3, 3].astype(np.float64)) @ T0[:3, :3].astype(np.float64)).astype(f32)
    nmp = len(pos)
    Ow0 = O.minus_Rt_t(T0[:3, :3], T0[:3, 3]).astype(np.float64)
    po = pos.astype(np.float64) - Ow0
    dist = np.linalg.norm(po, axis=1)
    nrm = po / dist[:, None] + rng.normal(0, 0.45, po.shape)
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(f32)
    octave = np.concatenate([k0[own, 5].astype(int), rng.integers(0, 8, n_extra)])
    max_d = (dist * scale[octave] * rng.uniform(0.8, 1.3, nmp)).astype(f32)
    min_d = (max_d / scale[-1]).astype(f32)
    bits = np.unpackbits(np.concatenate([d0[own], synth.descriptors(n_extra, seed)]), axis=1)
    for i in range(nmp):
        bits[i, rng.choice(256, rng.integers(0, 50), replace=False)] ^= 1
    mdesc = np.packbits(bits, axis=1)
    bad = rng.random(nmp) < 0.05
    nobs = np.concatenate([rng.integers(1, 6, len(own)), rng.integers(0, 4, n_extra)]).astype(np.int32)
    assoc0 = np.full(n0, -1, np.int32); assoc0[own] = np.arange(len(own))
    # view 1 keypoints that already hold a map point: extras on 20 %, and for some view 0 points the keypoint they project onto
    assoc1 = np.full(n1, -1, np.int32)
    holders = rng.permutation(n1)[:min(n1 // 5, n_extra)]
    assoc1[holders] = len(own) + np.arange(len(holders))
    for m in rng.permutation(len(own))[:120]:
        ux, uy = k0[own[m], 0] - dx, k0[own[m], 1] - dy
        j = int(np.argmin(np.abs(k1[:, 0] - ux) + np.abs(k1[:, 1] - uy)))
        if abs(k1[j, 0] - ux) + abs(k1[j, 1] - uy) < 2.5 and assoc1[j] < 0 and m not in assoc1:
            assoc1[j] = m
    outlier0 = rng.random(n0) < 0.05
    found = rng.random(nmp) < 0.1
    uright = [np.where(rng.random(len(k)) < 0.5, k[:, 0] - rng.uniform(2, 40, len(k)), -1).astype(f32) for k in (k0, k1)]
    # feature vectors: a feature of view 1 falls into the node of the view 0 feature it corresponds to 90 % of the time
    nodes = 50
    node0 = rng.integers(0, nodes, n0)
    near = np.array([int(np.argmin(np.abs(k0[:, 0] - dx - x) + np.abs(k0[:, 1] - dy - y))) for x, y in k1[:, :2]])
    node1 = np.where(rng.random(n1) < 0.9, node0[near], rng.integers(0, nodes, n1))
    fvs = []
    for node in (node0, node1):
        d = {}
        for i, k in enumerate(node):
            d.setdefault(int(k), []).append(i)
        fvs.append(sorted(d.items()))
    # tracking outputs of Frame::isInFrustum for view 1
    c1 = dict(cam, Rcw=T1[:3, :3], tcw=T1[:3, 3], Ow=O.minus_Rt_t(T1[:3, :3], T1[:3, 3]))
    fr = O.project_points(c1, pos, nrm, max_d, min_d, 0, 0.5, float(TH[0]))
    # fundamental matrix of a pure sideways motion along the image shift (the shifted image has no perspective change)
    tt = np.array([trel[0], trel[1], 0.0])
    Kin = np.linalg.inv(np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], np.float64))
    tx = np.array([[0, -tt[2], tt[1]], [tt[2], 0, -tt[0]], [-tt[1], tt[0], 0]])
    F12 = (Kin.T @ tx @ Kin).astype(f32)
    s = 1.3
    Scw = T1.astype(np.float64).copy(); Scw[:3, :] *= s
    Scw = Scw.astype(f32)
    sim = np.concatenate([[1.02], np.eye(3).ravel(), -trel]).astype(f32)  # p(view 0) = s12*R12*p(view 1) + t12
    return dict(cam=cam, T=[T0, T1], k=[k0, k1], d=[d0, d1], uright=uright, fv=fvs, pos=pos, nrm=nrm, max_d=max_d, min_d=min_d, mdesc=mdesc,
                bad=bad, nobs=nobs, assoc=[assoc0, assoc1], outlier0=outlier0, found=found, frustum=fr, F12=F12, Scw=Scw, sim=sim, scale=scale)


def oracle_results(sc):
    """What tests/cpp/guided_test.cc must print, computed with the scalar oracle."""
    cam, scale = sc["cam"], sc["scale"]
    T0, T1 = sc["T"]
    assoc0, assoc1 = sc["assoc"]
    pos, nrm, max_d, min_d, mdesc, bad, nobs = sc["pos"], sc["nrm"], sc["max_d"], sc["min_d"], sc["mdesc"], sc["bad"], sc["nobs"]
    nmp = len(pos)
    F = [O.OracleFrame(sc["k"][v], sc["d"][v], W, H) for v in range(2)]
    for v in range(2):
        F[v].scale, F[v].uright = scale, sc["uright"][v]
    res = []
    hold_obs = lambda a: (a >= 0) & (nobs[np.maximum(a, 0)] > 0)
    Ow = [O.minus_Rt_t(T[:3, :3], T[:3, 3]) for T in (T0, T1)]
    # 1. a-10
    fr = sc["frustum"]
    F[1].occupied = hold_obs(assoc1)
    mp = dict(in_view=fr["alive"].astype(bool), bad=bad, level=fr["level"], view_cos=fr["view_cos"], proj_x=fr["u"], proj_y=fr["v"], proj_xr=fr["ur"],
              desc=mdesc, has_obs=nobs > 0)
    nm, asg = O.search_by_projection_frame_mappoints(F[1], mp, float(TH[0]), 0.8)
    res += [[nm], np.where(asg >= 0, asg, assoc1)]
    # 2. a-11, stereo then mono
    alive = (assoc0 >= 0) & ~sc["outlier0"]
    pl = O.prologue_cur_last(cam, T1, pos[np.maximum(assoc0, 0)], alive)
    last = dict(pl, octave=sc["k"][0][:, 5].astype(np.int32), angle=sc["k"][0][:, 3], desc=mdesc[np.maximum(assoc0, 0)], has_obs=nobs[np.maximum(assoc0, 0)] > 0)
    for mono in (False, True):
        fw, bw = O.motion_direction(T1, T0, cam["mb"], mono)
        cleared = set()
        nm, asg = O.search_by_projection_cur_last(F[1], last, float(TH[1]), fw, bw, cleared=cleared)
        exp = np.where(asg >= 0, assoc0[np.maximum(asg, 0)], assoc1)
        exp[list(cleared)] = -1
        res += [[nm], exp]
    # 3. a-12 (Cur, KF)
    alive = (assoc0 >= 0) & ~bad[np.maximum(assoc0, 0)] & ~sc["found"][np.maximum(assoc0, 0)]
    idx = np.maximum(assoc0, 0)
    pr = O.prologue_scaled(cam, T1[:3, :3], T1[:3, 3], Ow[1], pos[idx], nrm[idx], max_d[idx], min_d[idx], alive, True, False, 0, "frame", False)
    F[1].occupied = assoc1 >= 0
    q = dict(pr, angle=sc["k"][0][:, 3], desc=mdesc[idx])
    cleared = set()
    nm, asg = O.search_by_projection_cur_kf(F[1], q, float(TH[2]), 64, cleared=cleared)
    exp = np.where(asg >= 0, assoc0[np.maximum(asg, 0)], assoc1)
    exp[list(cleared)] = -1
    res += [[nm], exp]
    # 4. a-12 (KF, Scw)
    R, t, ow = O.decompose_sim3(sc["Scw"])
    in_kf1 = np.zeros(nmp, bool); in_kf1[assoc1[assoc1 >= 0]] = True
    pr = O.prologue_scaled(cam, R, t, ow, pos, nrm, max_d, min_d, ~bad & ~in_kf1, False, True, 1, "image", True)
    nm, asg = O.search_by_projection_kf_sim3(F[1], assoc1 >= 0, dict(pr, desc=mdesc), int(TH[3]))
    res += [[nm], np.where(asg >= 0, asg, assoc1)]
    # 5. a-13 (KF, F)
    valid = (assoc0 >= 0) & ~bad[idx]
    nm, asg = O.search_by_bow_kf_f(sc["d"][0], sc["fv"][0], valid, sc["k"][0][:, 3], sc["d"][1], sc["fv"][1], sc["k"][1][:, 3], 0.75)
    res += [[nm], np.where(asg >= 0, assoc0[np.maximum(asg, 0)], -1)]
    # 6. a-14
    kf = [dict(desc=sc["d"][v], featvec=sc["fv"][v], has_mp=sc["assoc"][v] >= 0, uright=sc["uright"][v], x=sc["k"][v][:, 0], y=sc["k"][v][:, 1],
               angle=sc["k"][v][:, 3], octave=sc["k"][v][:, 5].astype(int)) for v in range(2)]
    c2 = O._mat3_vec(T1[:3, :3], Ow[0], T1[:3, 3])
    invz = f32(f32(1.0) / c2[2])
    epi = (f32(f32(f32(cam["fx"] * c2[0]) * invz) + cam["cx"]), f32(f32(f32(cam["fy"] * c2[1]) * invz) + cam["cy"]))
    sigma2 = (scale * scale).astype(f32)
    for only in (False, True):
        pairs = O.search_for_triangulation(kf[0], kf[1], sc["F12"], epi, scale, sigma2, only, check_ori=False)
        res += [[len(pairs)], np.array(pairs, np.int32).reshape(-1)]
    # 7. Fuse(KF, MPs)
    F[1].uright = sc["uright"][1]
    c1 = dict(cam, Rcw=T1[:3, :3], tcw=T1[:3, 3], Ow=Ow[1])
    pr = O.project_points(c1, pos, nrm, max_d, min_d, 1, 0.5, float(TH[4]))
    qq = dict(valid=pr["alive"].astype(bool) & ~bad & ~in_kf1, u=pr["u"], v=pr["v"], ur=pr["ur"], level=pr["level"], desc=mdesc)
    inv_sigma2 = (f32(1.0) / sigma2).astype(f32)
    _, best, _ = O.fuse_kf_mappoints(F[1], inv_sigma2, qq, float(TH[4]))
    held, nob, isbad, log, nf = assoc1.copy(), nobs.copy(), bad.copy(), [], 0
    for i in range(nmp):
        if best[i] < 0:
            continue
        other = held[best[i]]
        if other >= 0:
            if not isbad[other]:
                if nob[other] > nob[i]:
                    log += [2, i, other]; isbad[i] = True
                else:
                    log += [2, other, i]; isbad[other] = True
        else:
            log += [1, i, best[i]]; held[best[i]] = i; nob[i] += 1
        nf += 1
    res += [[nf], np.array(log, np.int32), held]
    # 8. Fuse(KF, Scw)
    found_kf = np.zeros(nmp, bool); found_kf[[m for m in assoc1 if m >= 0 and not bad[m]]] = True
    pr = O.prologue_scaled(cam, R, t, ow, pos, nrm, max_d, min_d, ~bad & ~found_kf, True, True, 1, "image", True)
    best = O.best_in_window(F[1], dict(pr, desc=mdesc), float(TH[5]), 50)
    held, repl, nf = assoc1.copy(), np.full(nmp, -1, np.int32), 0
    for i in range(nmp):
        if best[i] < 0:
            continue
        other = held[best[i]]
        if other >= 0:
            if not bad[other]:
                repl[i] = other
        else:
            held[best[i]] = i
        nf += 1
    res += [[nf], repl, held]
    # 9. SearchBySim3
    n0, n1 = len(assoc0), len(assoc1)
    m12 = np.full(n0, -1, np.int32)
    done1, done2 = np.zeros(n0, bool), np.zeros(n1, bool)
    where1 = {int(m): j for j, m in enumerate(assoc1) if m >= 0}
    for i in range(0, n0, 17):
        if assoc0[i] >= 0 and int(assoc0[i]) in where1:
            m12[i] = assoc0[i]; done1[i] = True; done2[where1[int(assoc0[i])]] = True
    s12, R12, t12 = f32(sc["sim"][0]), sc["sim"][1:10].reshape(3, 3), sc["sim"][10:13]
    sR12 = (R12 * s12).astype(f32)
    sR21 = (R12.T * f32(1.0 / float(s12))).astype(f32)
    t21 = -np.array(O._mat3_vec(sR21, t12, np.zeros(3, f32)), f32)
    i0, i1 = np.maximum(assoc0, 0), np.maximum(assoc1, 0)
    q1 = O.prologue_scaled(cam, sR21, t21, None, pos[i0], nrm[i0], max_d[i0], min_d[i0], (assoc0 >= 0) & ~done1 & ~bad[i0], True, True, 1, "image",
                           False, pre=(T0[:3, :3], T0[:3, 3]), dist_from_cam=True)
    q2 = O.prologue_scaled(cam, sR12, t12, None, pos[i1], nrm[i1], max_d[i1], min_d[i1], (assoc1 >= 0) & ~done2 & ~bad[i1], True, True, 1, "image",
                           False, pre=(T1[:3, :3], T1[:3, 3]), dist_from_cam=True)
    nf, mm = O.search_by_sim3(F[0], F[1], dict(q1, desc=mdesc[i0]), dict(q2, desc=mdesc[i1]), float(TH[6]))
    res += [[nf], np.where(mm >= 0, assoc1[np.maximum(mm, 0)], m12)]
    return [np.asarray(r, np.int32) for r in res]


NAMES = ["a10 n", "a10 map", "a11 stereo n", "a11 stereo map", "a11 mono n", "a11 mono map", "a12 cur-kf n", "a12 cur-kf map", "a12 kf-scw n",
         "a12 kf-scw map", "bow n", "bow map", "tri n", "tri pairs", "tri-stereo n", "tri-stereo pairs", "fuse n", "fuse log", "fuse map",
         "fuse-scw n", "fuse-scw replace", "fuse-scw map", "sim3 n", "sim3 map"]


