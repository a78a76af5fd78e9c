"""The remaining guided searches of ORBmatcher (SURVEY.md section 8 rows a-11 ... a-15) on arrays.

Same split as the other mirrors in matcher.py: the reference's pointer graph (Frame / KeyFrame / MapPoint) is flattened
by the caller into per-query arrays - the projection of each map point (u, v, ur), its predicted level, its descriptor
and the "is this query alive" gate the reference evaluates before the window lookup - and the device does the window
lookup (Frame::GetFeaturesInArea order) and every Hamming distance in one launch (DeviceFrameGrid.window_knn2 /
window_lists, orbm_list_distances). Searches whose candidate gate depends on matches made earlier in the same call
(vpMatched / mvpMapPoints state) replay the reference's ordered loop on the host over the device's candidate lists;
searches without such state take the device's (first minimum, best distance) directly.

Every function cites the reference lines it mirrors; `valid` is always "the reference reached the GetFeaturesInArea
call for this query" (map point exists, not bad, not already found, in front of the camera, inside the image and the
scale-invariance range, viewing angle gate - all caller-side geometry)."""
import numpy as np

from . import _lib
from .matcher import ORBmatcher, _desc, _p

f32 = np.float32


def _rot_bin(a1, a2, factor, L):
    r = f32(f32(a1) - f32(a2))
    if r < 0:
        r = f32(r + f32(360.0))
    b = int(np.floor(float(f32(r * factor)) + 0.5))  # C round() of a non-negative float
    return 0 if b == L else b


def _apply_rotation_consistency(self, rot, clear):
    """The 'Apply rotation consistency' epilogue shared by the searches (e.g. src/ORBmatcher.cc:1449-1469):
    entries of every histogram slot outside the three maxima are cleared; returns how many were removed."""
    keep = self.ComputeThreeMaxima([len(x) for x in rot])
    removed = 0
    for i in range(self.HISTO_LENGTH):
        if i in keep:
            continue
        for j in rot[i]:
            clear(j)
            removed += 1
    return removed


def _radius(th, scale_factors, level):
    return (f32(th) * np.asarray(scale_factors, f32)[np.asarray(level, np.int64)]).astype(f32)


# ---- a-11 ------------------------------------------------------------------------------------------------------
def _search_by_projection_cur_last(self, grid, kp_angle, uright, occupied, scale_factors, last, th, forward=False,
                                   backward=False):
    """ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)
    (/root/reference/src/ORBmatcher.cc:1330-1472). grid: DeviceFrameGrid of the current frame; kp_angle, uright:
    mvKeysUn[].angle / mvuRight of the current frame; occupied[kp]: the keypoint holds a map point with
    Observations() > 0 (1406-1408). last: per last-frame keypoint arrays valid, u, v, ur (= u - mbf*invzc), octave,
    angle, desc and has_obs (Observations() > 0 of its map point - what makes the keypoint it is assigned to occupied
    for later queries). forward / backward: bForward / bBackward (1350-1351), which pick the level window (1384-1389).
    Returns (nmatches, assigned[kp] = index into the last frame or -1)."""
    n = len(last["octave"])
    valid = np.asarray(last["valid"], bool)
    octv = np.asarray(last["octave"], np.int32)
    rad = _radius(th, scale_factors, octv)
    if forward:
        lo, hi = octv, np.full(n, -1, np.int32)
    elif backward:
        lo, hi = np.zeros(n, np.int32), octv
    else:
        lo, hi = octv - 1, octv + 1
    offsets, cands, dist = grid.window_lists(last["desc"], last["u"], last["v"], np.where(valid, rad, f32(0)).astype(f32), lo, hi)
    occupied = np.asarray(occupied, bool).copy()
    has_obs = np.asarray(last.get("has_obs", np.ones(n, bool)), bool)
    assigned = np.full(len(kp_angle), -1, np.int32)
    rot = [[] for _ in range(self.HISTO_LENGTH)]
    factor = f32(1.0) / f32(self.HISTO_LENGTH)
    nm = 0
    for i in range(n):
        if not valid[i] or offsets[i] == offsets[i + 1]:
            continue
        best, bi = 256, -1
        for k in range(offsets[i], offsets[i + 1]):
            i2 = int(cands[k])
            if occupied[i2]:
                continue
            if uright[i2] > 0 and abs(f32(last["ur"][i]) - f32(uright[i2])) > rad[i]:
                continue
            d = int(dist[k])
            if d < best:
                best, bi = d, i2
        if best <= self.TH_HIGH:
            assigned[bi] = i
            occupied[bi] = has_obs[i]
            nm += 1
            if self.mbCheckOrientation:
                rot[_rot_bin(last["angle"][i], kp_angle[bi], factor, self.HISTO_LENGTH)].append(bi)
    if self.mbCheckOrientation:
        def clear(j):
            assigned[j] = -1
        nm -= _apply_rotation_consistency(self, rot, clear)
    return nm, assigned


# ---- a-12 ------------------------------------------------------------------------------------------------------
def _search_by_projection_cur_kf(self, grid, kp_angle, occupied, scale_factors, q, th, orb_dist):
    """ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, ORBdist)
    (/root/reference/src/ORBmatcher.cc:1474-1601), relocalisation. occupied[kp]: CurrentFrame.mvpMapPoints[kp] is
    set (1544-1545). q: per keyframe map point arrays valid, u, v, level (nPredictedLevel), angle (pKF->mvKeysUn[i].angle),
    desc. Returns (nmatches, assigned[kp] = query index or -1)."""
    n = len(q["level"])
    valid = np.asarray(q["valid"], bool)
    lvl = np.asarray(q["level"], np.int32)
    rad = _radius(th, scale_factors, lvl)
    offsets, cands, dist = grid.window_lists(q["desc"], q["u"], q["v"], np.where(valid, rad, f32(0)).astype(f32), lvl - 1, lvl + 1)
    occupied = np.asarray(occupied, bool).copy()
    assigned = np.full(len(kp_angle), -1, np.int32)
    rot = [[] for _ in range(self.HISTO_LENGTH)]
    factor = f32(1.0) / f32(self.HISTO_LENGTH)
    nm = 0
    for i in range(n):
        if not valid[i] or offsets[i] == offsets[i + 1]:
            continue
        best, bi = 256, -1
        for k in range(offsets[i], offsets[i + 1]):
            i2 = int(cands[k])
            if occupied[i2]:
                continue
            d = int(dist[k])
            if d < best:
                best, bi = d, i2
        if best <= orb_dist:
            assigned[bi] = i
            occupied[bi] = True
            nm += 1
            if self.mbCheckOrientation:
                rot[_rot_bin(q["angle"][i], kp_angle[bi], factor, self.HISTO_LENGTH)].append(bi)
    if self.mbCheckOrientation:
        def clear(j):
            assigned[j] = -1
        nm -= _apply_rotation_consistency(self, rot, clear)
    return nm, assigned


def _search_by_projection_kf_sim3(self, grid, matched, scale_factors, q, th):
    """ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th)
    (/root/reference/src/ORBmatcher.cc:292-405), loop-closure guided search. matched[kp]: vpMatched[kp] is set on
    entry (372-373; the points themselves are excluded through `valid`, 319-320). q: valid, u, v, level, desc.
    Returns (nmatches, assigned[kp] = query index or -1 for the keypoints matched by this call)."""
    n = len(q["level"])
    valid = np.asarray(q["valid"], bool)
    lvl = np.asarray(q["level"], np.int32)
    rad = _radius(int(th), scale_factors, lvl)
    offsets, cands, dist = grid.window_lists(q["desc"], q["u"], q["v"], np.where(valid, rad, f32(0)).astype(f32), lvl - 1, lvl)
    matched = np.asarray(matched, bool).copy()
    assigned = np.full(len(matched), -1, np.int32)
    nm = 0
    for i in range(n):
        if not valid[i] or offsets[i] == offsets[i + 1]:
            continue
        best, bi = 256, -1
        for k in range(offsets[i], offsets[i + 1]):
            idx = int(cands[k])
            if matched[idx]:
                continue
            d = int(dist[k])
            if d < best:
                best, bi = d, idx
        if best <= self.TH_LOW:
            matched[bi] = True
            assigned[bi] = i
            nm += 1
    return nm, assigned


# ---- a-13 (KeyFrame, Frame) ----------------------------------------------------------------------------------------
def _merge_equal_nodes(fv1, fv2):
    """The lower_bound merge walk of two DBoW2 FeatureVectors (e.g. src/ORBmatcher.cc:182-266): pairs of index lists
    with the same node id, in node order."""
    out = []
    i = j = 0
    while i < len(fv1) and j < len(fv2):
        a, b = fv1[i][0], fv2[j][0]
        if a == b:
            out.append((fv1[i][1], fv2[j][1]))
            i += 1; j += 1
        elif a < b:
            i += 1
        else:
            j += 1
    return out


def _node_pair_distances(self, desc1, desc2, node_pairs):
    """All Hamming distances inside equal nodes in one launch: CSR over (idx1 occurrence) -> the idx2 list of its node."""
    q_rows, offsets, cands = [], [0], []
    for l1, l2 in node_pairs:
        for idx1 in l1:
            q_rows.append(idx1)
            cands.extend(l2)
            offsets.append(len(cands))
    q_rows = np.asarray(q_rows, np.int64)
    offsets = np.asarray(offsets, np.int32); cands = np.asarray(cands, np.int32)
    dist = np.empty(len(cands), np.int16)
    if len(cands):
        A = np.ascontiguousarray(desc1[q_rows])
        _lib.check(self._L.orbm_list_distances(self.device, _p(A), len(A), _p(desc2), len(desc2), _p(offsets), _p(cands), _p(dist)))
    return q_rows, offsets, cands, dist


def _search_by_bow_kf_f(self, desc_kf, featvec_kf, valid_kf, angle_kf, desc_f, featvec_f, angle_f):
    """ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, vpMapPointMatches) (/root/reference/src/ORBmatcher.cc:161-290).
    valid_kf[i]: the keyframe feature holds a map point that is not bad (193-199); angle_kf = pKF->mvKeysUn[].angle,
    angle_f = F.mvKeys[].angle (238-241). Returns (nmatches, assigned[f] = keyframe feature index or -1)."""
    desc_kf, desc_f = _desc(desc_kf), _desc(desc_f)
    q_rows, offsets, cands, dist = _node_pair_distances(self, desc_kf, desc_f, _merge_equal_nodes(featvec_kf, featvec_f))
    assigned = np.full(len(desc_f), -1, np.int32)
    rot = [[] for _ in range(self.HISTO_LENGTH)]
    factor = f32(1.0) / f32(self.HISTO_LENGTH)
    ratio = f32(self.mfNNratio)
    nm = 0
    for qi, ikf in enumerate(q_rows):
        if not valid_kf[ikf]:
            continue
        b1 = b2 = 256
        bi = -1
        for k in range(offsets[qi], offsets[qi + 1]):
            jf = int(cands[k])
            if assigned[jf] >= 0:
                continue
            d = int(dist[k])
            if d < b1:
                b2, b1, bi = b1, d, jf
            elif d < b2:
                b2 = d
        if b1 <= self.TH_LOW and f32(b1) < ratio * f32(b2):
            assigned[bi] = ikf
            if self.mbCheckOrientation:
                rot[_rot_bin(angle_kf[ikf], angle_f[bi], factor, self.HISTO_LENGTH)].append(bi)
            nm += 1
    if self.mbCheckOrientation:
        def clear(j):
            assigned[j] = -1
        nm -= _apply_rotation_consistency(self, rot, clear)
    return nm, assigned


# ---- a-14 ------------------------------------------------------------------------------------------------------
def _search_for_triangulation(self, kf1, kf2, F12, epipole, scale_factors2, level_sigma2_2, only_stereo=False):
    """ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)
    (/root/reference/src/ORBmatcher.cc:659-825) with CheckDistEpipolarLine (142-159). kf1 / kf2: dicts with desc,
    featvec, has_mp (GetMapPoint(idx) != NULL), uright, x, y, angle, octave (mvKeysUn). epipole = (ex, ey) (666-673).
    There is no match state between queries (vbMatched2 is never set in the reference), so every query is resolved
    from the device distances and vectorised gates. Returns the list of (idx1, idx2)."""
    d1, d2 = _desc(kf1["desc"]), _desc(kf2["desc"])
    q_rows, offsets, cands, dist = _node_pair_distances(self, d1, d2, _merge_equal_nodes(kf1["featvec"], kf2["featvec"]))
    F12 = np.asarray(F12, f32)
    ex, ey = f32(epipole[0]), f32(epipole[1])
    x1, y1 = np.asarray(kf1["x"], f32), np.asarray(kf1["y"], f32)
    x2, y2, o2 = np.asarray(kf2["x"], f32), np.asarray(kf2["y"], f32), np.asarray(kf2["octave"], np.int64)
    st1, st2 = np.asarray(kf1["uright"], f32) >= 0, np.asarray(kf2["uright"], f32) >= 0
    has1, has2 = np.asarray(kf1["has_mp"], bool), np.asarray(kf2["has_mp"], bool)
    sc2, sg2 = np.asarray(scale_factors2, f32), np.asarray(level_sigma2_2, f32)
    match12 = np.full(len(d1), -1, np.int32)
    rot = [[] for _ in range(self.HISTO_LENGTH)]
    factor = f32(1.0) / f32(self.HISTO_LENGTH)
    nm = 0
    for qi, i1 in enumerate(q_rows):
        if has1[i1] or (only_stereo and not st1[i1]):
            continue
        c = cands[offsets[qi]:offsets[qi + 1]].astype(np.int64)
        if not len(c):
            continue
        dd = dist[offsets[qi]:offsets[qi + 1]].astype(np.int32)
        ok = ~has2[c]
        if only_stereo:
            ok &= st2[c]
        ok &= dd <= self.TH_LOW
        if not st1[i1]:  # epipole distance gate, only when neither side is stereo (731-737)
            dx, dy = (ex - x2[c]).astype(f32), (ey - y2[c]).astype(f32)
            near = (dx * dx + dy * dy).astype(f32) < (f32(100) * sc2[o2[c]]).astype(f32)
            ok &= ~(near & ~st2[c])
        # epipolar line of kp1 in image 2 (145-158), float, left to right
        a = f32(f32(x1[i1] * F12[0, 0] + y1[i1] * F12[1, 0]) + F12[2, 0])
        b = f32(f32(x1[i1] * F12[0, 1] + y1[i1] * F12[1, 1]) + F12[2, 1])
        cc = f32(f32(x1[i1] * F12[0, 2] + y1[i1] * F12[1, 2]) + F12[2, 2])
        den = f32(a * a + b * b)
        if den == 0:
            continue
        num = ((a * x2[c]).astype(f32) + (b * y2[c]).astype(f32)).astype(f32) + cc
        dsqr = ((num * num).astype(f32) / den).astype(f32)
        ok &= dsqr.astype(np.float64) < 3.84 * sg2[o2[c]].astype(np.float64)
        # running 'dist > bestDist -> continue' (727) with bestDist only lowered by accepted candidates: the winner is
        # the last accepted candidate whose distance is <= every accepted distance before it = the last minimum.
        if not ok.any():
            continue
        dk = np.where(ok, dd, 1 << 20)
        m = dk.min()
        best = int(c[len(dk) - 1 - int(np.argmin(dk[::-1]))])
        assert m <= self.TH_LOW
        match12[i1] = best
        nm += 1
        if self.mbCheckOrientation:
            rot[_rot_bin(kf1["angle"][i1], kf2["angle"][best], factor, self.HISTO_LENGTH)].append(int(i1))
    if self.mbCheckOrientation:
        def clear(j):
            match12[j] = -1
        nm -= _apply_rotation_consistency(self, rot, clear)
    return [(int(i), int(match12[i])) for i in range(len(match12)) if match12[i] >= 0]


# ---- a-15 ------------------------------------------------------------------------------------------------------
def _fuse_kf_mappoints(self, grid, kp_x, kp_y, kp_octave, uright, inv_level_sigma2, scale_factors, q, th=3.0):
    """The search of ORBmatcher::Fuse(KeyFrame *pKF, vpMapPoints, th) (/root/reference/src/ORBmatcher.cc:827-977):
    for every map point the keypoint it fuses into (or -1). q: valid, u, v, ur, level, desc. The replace / add
    bookkeeping on the graph (953-972) stays with the caller, which owns the MapPoint objects.
    Returns (nFused, best_idx[query], best_dist[query])."""
    n = len(q["level"])
    valid = np.asarray(q["valid"], bool)
    lvl = np.asarray(q["level"], np.int32)
    rad = _radius(th, scale_factors, lvl)
    offsets, cands, dist = grid.window_lists(q["desc"], q["u"], q["v"], np.where(valid, rad, f32(0)).astype(f32), lvl - 1, lvl)
    kp_x, kp_y, uright = np.asarray(kp_x, f32), np.asarray(kp_y, f32), np.asarray(uright, f32)
    inv_s2 = np.asarray(inv_level_sigma2, f32)[np.asarray(kp_octave, np.int64)]
    best_idx = np.full(n, -1, np.int32)
    best_dist = np.full(n, 256, np.int32)
    for i in range(n):
        s, e = offsets[i], offsets[i + 1]
        if not valid[i] or s == e:
            continue
        c = cands[s:e].astype(np.int64)
        ex = (f32(q["u"][i]) - kp_x[c]).astype(f32)
        ey = (f32(q["v"][i]) - kp_y[c]).astype(f32)
        er = (f32(q["ur"][i]) - uright[c]).astype(f32)
        e2m = ((ex * ex).astype(f32) + (ey * ey).astype(f32)).astype(f32)
        e2s = (e2m + (er * er).astype(f32)).astype(f32)
        stereo = uright[c] >= 0
        chi = np.where(stereo, e2s, e2m) * inv_s2[c]  # float product, compared as double with 7.8 / 5.99 (922, 934)
        ok = chi.astype(f32).astype(np.float64) <= np.where(stereo, 7.8, 5.99)
        if not ok.any():
            continue
        dk = np.where(ok, dist[s:e].astype(np.int32), 1 << 20)
        k = int(np.argmin(dk))  # first minimum = strict '<' update (942)
        best_dist[i] = dk[k]
        if dk[k] <= self.TH_LOW:
            best_idx[i] = c[k]
    return int((best_idx >= 0).sum()), best_idx, best_dist


def _best_in_window(self, grid, scale_factors, q, th, limit):
    """Stateless guided search on the device: first minimum inside the window at levels [pred-1, pred]."""
    valid = np.asarray(q["valid"], bool)
    lvl = np.asarray(q["level"], np.int32)
    rad = _radius(th, scale_factors, lvl)
    idx, d1, _, _, _ = grid.window_knn2(q["desc"], q["u"], q["v"], np.where(valid, rad, f32(0)).astype(f32), lvl - 1, lvl)
    hit = valid & (idx >= 0) & (d1 <= limit)
    return np.where(hit, idx, -1).astype(np.int32), np.where(valid & (idx >= 0), d1, 256).astype(np.int32)


def _fuse_kf_sim3(self, grid, scale_factors, q, th):
    """The search of ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, vpPoints, th, vpReplacePoint)
    (/root/reference/src/ORBmatcher.cc:979-1102), entirely on the device. Returns (nFused, best_idx, best_dist)."""
    bi, bd = _best_in_window(self, grid, scale_factors, q, th, self.TH_LOW)
    return int((bi >= 0).sum()), bi, bd


def _search_by_sim3(self, grid1, grid2, scale_factors1, scale_factors2, q1, q2, th):
    """ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (/root/reference/src/ORBmatcher.cc:1104-1328).
    q1: the map points of KF1 projected into KF2 (valid excludes vbAlreadyMatched1, 1161-1162), q2: those of KF2 into KF1.
    Two device searches and the mutual-agreement check (1311-1325). Returns (nFound, match12[i1] = idx2 or -1)."""
    m1, _ = _best_in_window(self, grid2, scale_factors2, q1, th, self.TH_HIGH)
    m2, _ = _best_in_window(self, grid1, scale_factors1, q2, th, self.TH_HIGH)
    match12 = np.full(len(m1), -1, np.int32)
    for i1 in range(len(m1)):
        i2 = m1[i1]
        if i2 >= 0 and i2 < len(m2) and m2[i2] == i1:
            match12[i1] = i2
    return int((match12 >= 0).sum()), match12


ORBmatcher.SearchByProjection_Cur_Last = _search_by_projection_cur_last
ORBmatcher.SearchByProjection_Cur_KF = _search_by_projection_cur_kf
ORBmatcher.SearchByProjection_KF_Sim3 = _search_by_projection_kf_sim3
ORBmatcher.SearchByBoW_KF_F = _search_by_bow_kf_f
ORBmatcher.SearchForTriangulation = _search_for_triangulation
ORBmatcher.Fuse_KF_MapPoints = _fuse_kf_mappoints
ORBmatcher.Fuse_KF_Sim3 = _fuse_kf_sim3
ORBmatcher.SearchBySim3 = _search_by_sim3
