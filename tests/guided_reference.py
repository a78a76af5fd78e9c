"""The guided-search scenario of tests/guided_scenario.py run through the reference's OWN ORBmatcher / Frame / KeyFrame /
MapPoint (oracle/_ref/libref_slam.so via oracle/ref_slam.py). Shared by tests/test_oracle_vs_reference_matcher.py and
tools/make_golden.py (which freezes these results into tests/golden for the GPU box)."""
import numpy as np

import guided_scenario as G
import ref_slam as R

f32 = np.float32
W, H = G.W, G.H


class RefScene:
    """The scenario of guided_scenario.make_scenario as the reference's own objects: two Frames / KeyFrames with poses,
    associations and feature vectors, and the map points with their descriptors, normals, distance ranges, observation
    counts and bad flags."""

    def __init__(self, sc, kf_views=(0, 1)):
        cam = sc["cam"]
        self.sc = sc
        w = self.w = R.World()
        K4 = [cam["fx"], cam["fy"], cam["cx"], cam["cy"]]
        self.f = [w.frame_arrays(sc["k"][v], sc["d"][v], K4, W, H, uright=sc["uright"][v], bf=float(cam["bf"]), Tcw=sc["T"][v]) for v in range(2)]
        self.kf = [w.keyframe(self.f[v]) if v in kf_views else None for v in range(2)]
        ref = self.kf[kf_views[0]]
        nmp = len(sc["pos"])
        for i in range(nmp):
            m = w.mappoint(sc["pos"][i], ref)
            assert m == i
            w.mp_set(m, desc=sc["mdesc"][i], normal=sc["nrm"][i], min_max=(sc["min_d"][i], sc["max_d"][i]), n_obs=sc["nobs"][i], bad=sc["bad"][i])
        for v in range(2):
            w.frame_set_featvec(self.f[v], sc["fv"][v])
            if self.kf[v] is not None:
                w.kf_set_featvec(self.kf[v], sc["fv"][v])

    def hold(self, v, frame=True, kf=True, observe=False):
        """view v's keypoints hold the scenario's associations (Frame::mvpMapPoints / KeyFrame::mvpMapPoints)."""
        a = self.sc["assoc"][v]
        if frame:
            self.w.frame_set_mappoints(self.f[v], a)
        if kf and self.kf[v] is not None:
            self.w.kf_set_mappoints(self.kf[v], a)
            if observe:
                for idx in np.flatnonzero(a >= 0):
                    self.w.mp_set_observation(int(a[idx]), self.kf[v], int(idx))


def reference_results(sc):
    """The list guided_scenario.oracle_results produces, computed by the reference's ORBmatcher on the reference's objects."""
    TH = G.TH
    nmp = len(sc["pos"])
    assoc0, assoc1 = sc["assoc"]
    res = []
    # 1. a-10: Frame::isInFrustum for every point, then SearchByProjection(F, vpMapPoints, th)
    s = RefScene(sc); s.hold(1)
    fr = np.stack([s.w.is_in_frustum(s.f[1], m, 0.5) for m in range(nmp)])
    nm = s.w.search_by_projection_local(s.f[1], np.arange(nmp), float(TH[0]), 0.8)
    res += [[nm], s.w.frame_get_mappoints(s.f[1])]
    frustum = fr
    # 2. a-11 stereo / mono
    for mono in (False, True):
        s = RefScene(sc); s.hold(0); s.hold(1)
        s.w.frame_set_outliers(s.f[0], sc["outlier0"])
        nm = s.w.search_by_projection_last(s.f[1], s.f[0], float(TH[1]), mono, 0.9, True)
        res += [[nm], s.w.frame_get_mappoints(s.f[1])]
    # 3. a-12 (Cur, KF, sAlreadyFound, th, ORBdist)
    s = RefScene(sc); s.hold(0); s.hold(1)
    nm = s.w.search_by_projection_kf(s.f[1], s.kf[0], np.flatnonzero(sc["found"]), float(TH[2]), 64, 0.9, True)
    res += [[nm], s.w.frame_get_mappoints(s.f[1])]
    # 4. a-12 (KF, Scw, vpPoints, vpMatched, th)
    s = RefScene(sc); s.hold(1)
    nm, matched = s.w.search_by_projection_sim3(s.kf[1], sc["Scw"], np.arange(nmp), assoc1, int(TH[3]))
    res += [[nm], matched]
    # 5. a-13 (KF, F)
    s = RefScene(sc); s.hold(0)
    nm, out = s.w.search_by_bow_kf_f(s.kf[0], s.f[1], 0.75, True)
    res += [[nm], out]
    # 6. a-14
    for only in (False, True):
        s = RefScene(sc); s.hold(0); s.hold(1)
        pairs = s.w.search_for_triangulation(s.kf[0], s.kf[1], sc["F12"], only, 0.6, False)
        res += [[len(pairs)], np.array(pairs, np.int32).reshape(-1)]
    # 7. Fuse(KF, vpMapPoints, th): compared through the graph state it leaves (see test below)
    s = RefScene(sc); s.hold(1, observe=True)
    nf = s.w.fuse(s.kf[1], np.arange(nmp), float(TH[4]))
    state = [s.w.mp_get(m) for m in range(nmp)]
    fuse = dict(n=nf, held=s.w.kf_get_mappoints(s.kf[1], len(assoc1)), bad=np.array([x["bad"] for x in state]),
                nobs=np.array([x["n_obs"] for x in state]), replaced=np.array([s.w.mp_replaced(m) for m in range(nmp)]))
    # 8. Fuse(KF, Scw, vpPoints, th, vpReplacePoint)
    s = RefScene(sc); s.hold(1, observe=True)
    nf, repl = s.w.fuse_sim3(s.kf[1], sc["Scw"], np.arange(nmp), float(TH[5]))
    res8 = [[nf], repl, s.w.kf_get_mappoints(s.kf[1], len(assoc1))]
    # 9. SearchBySim3 (vbAlreadyMatched2 comes from MapPoint::GetIndexInKeyFrame: the points need their observations)
    s = RefScene(sc); s.hold(0); s.hold(1, observe=True)
    n0 = len(assoc0)
    m12 = np.full(n0, -1, np.int32)
    where1 = {int(m): j for j, m in enumerate(assoc1) if m >= 0}
    for i in range(0, n0, 17):
        if assoc0[i] >= 0 and int(assoc0[i]) in where1:
            m12[i] = assoc0[i]
    sim = sc["sim"]
    nf, mm = s.w.search_by_sim3(s.kf[0], s.kf[1], m12, float(sim[0]), sim[1:10], sim[10:13], float(TH[6]))
    res9 = [[nf], mm]
    return [np.asarray(r, np.int32) for r in res], fuse, [np.asarray(r, np.int32) for r in res8 + res9], frustum


def fuse_graph_model(sc, best):
    """What the reference's graph calls inside Fuse(KF, vpMapPoints, th) (src/ORBmatcher.cc:950-969) do to the scene of
    RefScene.hold(1, observe=True), where every map point is observed by at most the target keyframe:
    MapPoint::AddObservation (+2 observations on a stereo keypoint, src/MapPoint.cc:102-113), KeyFrame::AddMapPoint,
    MapPoint::Replace (src/MapPoint.cc:181-221: the replaced point turns bad and hands its keyframe slot over)."""
    assoc1 = sc["assoc"][1]
    held, nob, bad = assoc1.copy(), sc["nobs"].copy(), sc["bad"].copy()
    nmp = len(sc["pos"])
    replaced = np.full(nmp, -1, np.int32)
    observed_at = {int(m): int(j) for j, m in enumerate(assoc1) if m >= 0}   # mObservations[kf1]
    stereo = sc["uright"][1] >= 0
    nf = 0

    def replace(this, by):   # this->Replace(by)
        bad[this] = True
        replaced[this] = by
        if this in observed_at:
            j = observed_at.pop(this)
            if by not in observed_at:
                held[j] = by
                observed_at[by] = j
                nob[by] += 2 if stereo[j] else 1
            else:
                held[j] = -1

    for i in range(nmp):
        if best[i] < 0:
            continue
        j = int(best[i])
        other = int(held[j])
        if other >= 0:
            if not bad[other]:
                if nob[other] > nob[i]:
                    replace(i, other)
                else:
                    replace(other, i)
        else:
            observed_at[i] = j
            nob[i] += 2 if stereo[j] else 1
            held[j] = i
        nf += 1
    return dict(n=nf, held=held, bad=bad, nobs=nob, replaced=replaced)


