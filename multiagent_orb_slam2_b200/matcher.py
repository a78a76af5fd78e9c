"""Host-side mirror of ORB_SLAM2::ORBmatcher (/root/reference/include/ORBmatcher.h:37-102) over the
C ABI, on plain arrays: the reference's searches walk Frame/KeyFrame/MapPoint pointer graphs (out of
scope); what they compute on descriptors is exposed here with the same
constants, thresholds and acceptance rules."""
import ctypes as C

import numpy as np

from . import _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _desc(a):
    a = np.ascontiguousarray(a, np.uint8)
    assert a.ndim == 2 and a.shape[1] == 32
    return a


class ORBmatcher:
    TH_HIGH = 100      # src/ORBmatcher.cc:37
    TH_LOW = 50        # :38
    HISTO_LENGTH = 30  # :39

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio = float(np.float32(nnratio))
        self.mbCheckOrientation = bool(checkOri)
        self.device = int(device)
        self._L = _lib.lib()

    @staticmethod
    def DescriptorDistance(a, b):
        """Host inline, as in the reference (src/ORBmatcher.cc:1649-1665)."""
        a = np.ascontiguousarray(a, np.uint8).reshape(32); b = np.ascontiguousarray(b, np.uint8).reshape(32)
        return int(np.unpackbits(a ^ b).sum())

    # -- device primitives on host arrays -----------------------------------------------------------
    def knn2(self, A, B):
        """Best / second-best Hamming distance and first-minimum index of every row of A in B."""
        A, B = _desc(A), _desc(B)
        idx = np.empty(len(A), np.int32); d1 = np.empty(len(A), np.int32); d2 = np.empty(len(A), np.int32)
        _lib.check(self._L.orbm_knn2(self.device, _p(A), len(A), _p(B), len(B), _p(idx), _p(d1), _p(d2)))
        return idx, d1, d2

    def knn2_lists(self, A, B, offsets, cands):
        A, B = _desc(A), _desc(B)
        offsets = np.ascontiguousarray(offsets, np.int32); cands = np.ascontiguousarray(cands, np.int32)
        assert len(offsets) == len(A) + 1
        idx = np.empty(len(A), np.int32); d1 = np.empty(len(A), np.int32); d2 = np.empty(len(A), np.int32)
        _lib.check(self._L.orbm_knn2_lists(self.device, _p(A), len(A), _p(B), len(B), _p(offsets), _p(cands), _p(idx), _p(d1), _p(d2)))
        return idx, d1, d2

    def distance_matrix(self, A, B):
        A, B = _desc(A), _desc(B)
        out = np.empty((len(A), len(B)), np.int16)
        _lib.check(self._L.orbm_distance_matrix(self.device, _p(A), len(A), _p(B), len(B), _p(out)))
        return out

    # -- acceptance rules ---------------------------------------------------------------------------
    def accept(self, idx, d1, d2, th=None, inclusive=False):
        """match[i] = idx[i] if d1 (< | <=) th and float(d1) < nnratio*float(d2) else -1
        (src/ORBmatcher.cc:600-603 strict; 229-232 / 461-463 inclusive)."""
        th = self.TH_LOW if th is None else th
        ok = (d1 <= th) if inclusive else (d1 < th)
        ok &= d1.astype(np.float32) < np.float32(self.mfNNratio) * d2.astype(np.float32)
        ok &= idx >= 0
        return np.where(ok, idx, -1).astype(np.int32)

    def SearchBruteForce(self, A, B, th=None):
        """The SearchByBoW(KF,KF) inner loop with the vocabulary gate removed (BASELINE configs 3, 5):
        ratio-tested nearest neighbour of every row of A in B. Returns match indices (-1 = none)."""
        idx, d1, d2 = self.knn2(A, B)
        return self.accept(idx, d1, d2, th)

    @staticmethod
    def ComputeThreeMaxima(hist_sizes):
        """src/ORBmatcher.cc:1603-1644 on the bin populations; returns (ind1, ind2, ind3)."""
        max1 = max2 = max3 = 0
        i1 = i2 = i3 = -1
        for i, s in enumerate(hist_sizes):
            s = int(s)
            if s > max1:
                max3, max2, max1 = max2, max1, s
                i3, i2, i1 = i2, i1, i
            elif s > max2:
                max3, max2 = max2, s
                i3, i2 = i2, i
            elif s > max3:
                max3, i3 = s, i
        lim = np.float32(0.1) * np.float32(max1)  # 0.1f*(float)max1, evaluated in fp32
        if np.float32(max2) < lim:
            i2 = i3 = -1
        elif np.float32(max3) < lim:
            i3 = -1
        return i1, i2, i3


class FrameGrid:
    """Host mirror of the reference Frame's 64 x 48 lookup grid: AssignFeaturesToGrid
    (/root/reference/src/Frame.cc:230-245), PosInGrid (382-392), GetFeaturesInArea (327-380).
    Keypoints as the KP_DTYPE structured array of ORBextractor; candidates come back in the
    reference's order (ix-major, iy, insertion)."""
    COLS, ROWS = 64, 48

    def __init__(self, kps, desc, width, height, min_x=0.0, min_y=0.0):
        f32 = np.float32
        self.kps, self.desc = kps, np.ascontiguousarray(desc, np.uint8)
        self.x, self.y, self.octave = kps["x"].astype(f32), kps["y"].astype(f32), kps["octave"].astype(np.int32)
        self.angle = kps["angle"].astype(f32)
        self.min_x, self.min_y = f32(min_x), f32(min_y)
        self.w_inv = f32(self.COLS) / f32(f32(width) - self.min_x)
        self.h_inv = f32(self.ROWS) / f32(f32(height) - self.min_y)
        fx = (self.x - self.min_x) * self.w_inv
        fy = (self.y - self.min_y) * self.h_inv
        px = (np.sign(fx) * np.floor(np.abs(fx) + f32(0.5))).astype(np.int64)  # C round(): half away from zero
        py = (np.sign(fy) * np.floor(np.abs(fy) + f32(0.5))).astype(np.int64)
        self.cells = {}
        for i in np.flatnonzero((px >= 0) & (px < self.COLS) & (py >= 0) & (py < self.ROWS)):
            self.cells.setdefault((int(px[i]), int(py[i])), []).append(int(i))

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        f32 = np.float32
        x, y, r = f32(x), f32(y), f32(r)
        c0 = max(0, int(np.floor((x - self.min_x - r) * self.w_inv)))
        c1 = min(self.COLS - 1, int(np.ceil((x - self.min_x + r) * self.w_inv)))
        r0 = max(0, int(np.floor((y - self.min_y - r) * self.h_inv)))
        r1 = min(self.ROWS - 1, int(np.ceil((y - self.min_y + r) * self.h_inv)))
        if c0 >= self.COLS or c1 < 0 or r0 >= self.ROWS or r1 < 0:
            return []
        check = minLevel > 0 or maxLevel >= 0
        out = []
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for j in self.cells.get((ix, iy), ()):
                    if check and (self.octave[j] < minLevel or (maxLevel >= 0 and self.octave[j] > maxLevel)):
                        continue
                    if abs(self.x[j] - x) < r and abs(self.y[j] - y) < r:
                        out.append(j)
        return out


def _search_for_initialization(self, F1, F2, vbPrevMatched, windowSize=10, device_grid=None):
    """ORBmatcher::SearchForInitialization (/root/reference/src/ORBmatcher.cc:407-522) on two FrameGrid
    objects. Candidate gating in the reference's order - on the host (F2.GetFeaturesInArea) or, with
    device_grid (a DeviceFrameGrid of frame 2), on the device (orbm_window_lists_device); device: all Hamming
    distances of the gated pairs; host: the reference's ordered, stateful resolve replayed over the precomputed
    distances. vbPrevMatched: (n1, 2) float32, updated in place. Returns (nmatches, vnMatches12)."""
    INT_MAX = 2**31 - 1
    n1, n2 = len(F1.kps), len(F2.kps)
    if device_grid is not None:
        r = np.where(F1.octave <= 0, np.float32(windowSize), np.float32(0)).astype(np.float32)  # level1 > 0: no window
        offsets, cands, dist = device_grid.window_lists(F1.desc, vbPrevMatched[:, 0], vbPrevMatched[:, 1], r, 0, 0)
    else:
        offsets = np.zeros(n1 + 1, np.int32)
        cands = []
        for i1 in range(n1):
            if F1.octave[i1] <= 0:
                cands += F2.GetFeaturesInArea(vbPrevMatched[i1, 0], vbPrevMatched[i1, 1], windowSize, 0, 0)
            offsets[i1 + 1] = len(cands)
        cands = np.asarray(cands, np.int32)
        dist = np.empty(len(cands), np.int16)
        if len(cands):
            _lib.check(self._L.orbm_list_distances(self.device, _p(F1.desc), n1, _p(F2.desc), n2, _p(offsets), _p(cands), _p(dist)))
    m12 = np.full(n1, -1, np.int32)
    m21 = np.full(n2, -1, np.int32)
    mdist = np.full(n2, INT_MAX, np.int64)
    rot = [[] for _ in range(self.HISTO_LENGTH)]
    factor = np.float32(1.0) / np.float32(self.HISTO_LENGTH)
    ratio = np.float32(self.mfNNratio)
    nm = 0
    for i1 in range(n1):
        lo, hi = offsets[i1], offsets[i1 + 1]
        if lo == hi:
            continue
        best = best2 = INT_MAX
        bidx = -1
        for k in range(lo, hi):
            i2, d = int(cands[k]), int(dist[k])
            if mdist[i2] <= d:
                continue
            if d < best:
                best2, best, bidx = best, d, i2
            elif d < best2:
                best2 = d
        if best <= self.TH_LOW and np.float32(best) < np.float32(best2) * ratio:
            if m21[bidx] >= 0:
                m12[m21[bidx]] = -1
                nm -= 1
            m12[i1], m21[bidx], mdist[bidx] = bidx, i1, best
            nm += 1
            if self.mbCheckOrientation:
                r = np.float32(F1.angle[i1] - F2.angle[bidx])
                if r < 0:
                    r = np.float32(r + np.float32(360.0))
                b = int(np.floor(float(np.float32(r * factor)) + 0.5))
                rot[0 if b == self.HISTO_LENGTH else b].append(i1)
    if self.mbCheckOrientation:
        keep = self.ComputeThreeMaxima([len(x) for x in rot])
        for i in range(self.HISTO_LENGTH):
            if i not in keep:
                for idx1 in rot[i]:
                    if m12[idx1] >= 0:
                        m12[idx1] = -1
                        nm -= 1
    ok = m12 >= 0
    vbPrevMatched[ok, 0] = F2.x[m12[ok]]
    vbPrevMatched[ok, 1] = F2.y[m12[ok]]
    return nm, m12


ORBmatcher.SearchForInitialization = _search_for_initialization


def _search_by_bow_kf_kf(self, desc1, featvec1, valid1, angle1, desc2, featvec2, valid2, angle2):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (/root/reference/src/ORBmatcher.cc:524-657),
    the matcher MapFusion::ComputeSim3 and CovisibilityDiscovery call (src/MapFusion.cc:275, 849).
    featvec*: DBoW2 FeatureVector as a list of (node_id, [feature indices]) sorted by node id;
    valid*: per feature, "has a map point that is not bad" (the pMP / isBad() tests at 560-564, 576-582).
    Device: all Hamming distances inside equal nodes in one launch; host: the reference's ordered,
    stateful (vbMatched2) resolve. Returns (nmatches, match12) with match12[i1] = i2 or -1."""
    desc1, desc2 = _desc(desc1), _desc(desc2)
    n1, n2 = len(desc1), len(desc2)
    # merge walk of the two feature vectors (552-634): pairs of index lists with equal node id
    node_pairs = []
    i = j = 0
    while i < len(featvec1) and j < len(featvec2):
        a, b = featvec1[i][0], featvec2[j][0]
        if a == b:
            node_pairs.append((featvec1[i][1], featvec2[j][1]))
            i += 1; j += 1
        elif a < b:
            i += 1  # lower_bound(f2it->first): first element not less than b
        else:
            j += 1
    # CSR over (idx1 occurrence) -> candidates idx2 of the same node, in the reference's loop order
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
        _lib.check(self._L.orbm_list_distances(self.device, _p(A), len(A), _p(desc2), n2, _p(offsets), _p(cands), _p(dist)))
    match12 = np.full(n1, -1, np.int32)
    matched2 = np.zeros(n2, bool)
    rot = [[] for _ in range(self.HISTO_LENGTH)]
    factor = np.float32(1.0) / np.float32(self.HISTO_LENGTH)
    ratio = np.float32(self.mfNNratio)
    nm = 0
    for q, idx1 in enumerate(q_rows):
        if not valid1[idx1]:
            continue
        b1 = b2 = 256
        bidx = -1
        for k in range(offsets[q], offsets[q + 1]):
            idx2 = int(cands[k])
            if matched2[idx2] or not valid2[idx2]:
                continue
            d = int(dist[k])
            if d < b1:
                b2, b1, bidx = b1, d, idx2
            elif d < b2:
                b2 = d
        if b1 < self.TH_LOW and np.float32(b1) < ratio * np.float32(b2):
            match12[idx1] = bidx
            matched2[bidx] = True
            if self.mbCheckOrientation:
                r = np.float32(np.float32(angle1[idx1]) - np.float32(angle2[bidx]))
                if r < 0:
                    r = np.float32(r + np.float32(360.0))
                b = int(np.floor(float(np.float32(r * factor)) + 0.5))
                rot[0 if b == self.HISTO_LENGTH else b].append(int(idx1))
            nm += 1
    if self.mbCheckOrientation:
        keep = self.ComputeThreeMaxima([len(x) for x in rot])
        for i in range(self.HISTO_LENGTH):
            if i not in keep:
                for idx1 in rot[i]:
                    match12[idx1] = -1
                    nm -= 1
    return nm, match12


def _distinctive_descriptor(self, descs):
    """MapPoint::ComputeDistinctiveDescriptors (/root/reference/src/MapPoint.cc:246-311): index of the
    observation with the least median Hamming distance to the others (first one wins ties). The N x N
    distances come from the device (orbm_distance_matrix)."""
    descs = _desc(descs)
    n = len(descs)
    if n == 0:
        return -1
    D = self.distance_matrix(descs, descs).astype(np.int32)
    med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
    return int(np.argmin(med))  # argmin returns the first minimum, like the strict '<' loop


ORBmatcher.SearchByBoW_KF_KF = _search_by_bow_kf_kf
ORBmatcher.ComputeDistinctiveDescriptor = _distinctive_descriptor


def _search_by_projection_frame_mappoints(self, grid, kp_octave, uright, occupied, scale_factors, mp, th):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (/root/reference/src/ORBmatcher.cc:45-131)
    on arrays, with the candidate gate and all Hamming distances on the device (DeviceFrameGrid.window_lists) and the
    reference's ordered, stateful loop replayed on the host. mp: per-map-point arrays in_view (mbTrackInView), bad,
    level (mnTrackScaleLevel), view_cos, proj_x, proj_y, proj_xr, desc. occupied[kp]: the keypoint already holds a map
    point with observations (90-92). Returns (nmatches, assigned[kp] = map point index or -1)."""
    f32 = np.float32
    n_mp = len(mp["level"])
    live = np.asarray(mp["in_view"], bool) & ~np.asarray(mp["bad"], bool)
    lvl = np.asarray(mp["level"], np.int32)
    r = np.where(np.asarray(mp["view_cos"], f32) > f32(0.998), f32(2.5), f32(4.0)).astype(f32)  # RadiusByViewingCos (133-139)
    if th != 1.0:
        r = (r * f32(th)).astype(f32)
    rad = (r * np.asarray(scale_factors, f32)[lvl]).astype(f32)
    rad_q = np.where(live, rad, f32(0)).astype(f32)  # dead map points: empty window
    offsets, cands, dist = grid.window_lists(mp["desc"], mp["proj_x"], mp["proj_y"], rad_q, lvl - 1, lvl)
    return _replay_frame_mappoints(self, offsets, cands, dist, live, rad, mp["proj_xr"], kp_octave, uright, occupied, mp.get("has_obs"))


def _replay_frame_mappoints(self, offsets, cands, dist, live, rad, proj_xr, kp_octave, uright, occupied, has_obs=None):
    """The ordered loop of src/ORBmatcher.cc:52-128 over the device's candidate lists."""
    f32 = np.float32
    occupied = np.asarray(occupied, bool).copy()
    assigned = np.full(len(kp_octave), -1, np.int32)
    ratio = f32(self.mfNNratio)
    nm = 0
    for i in range(len(live)):
        if not live[i] or offsets[i] == offsets[i + 1]:
            continue
        b1 = b2 = 256
        l1 = l2 = bi = -1
        for k in range(offsets[i], offsets[i + 1]):
            idx = int(cands[k])
            if occupied[idx]:
                continue
            if uright[idx] > 0 and abs(f32(proj_xr[i]) - f32(uright[idx])) > rad[i]:
                continue
            d = int(dist[k])
            if d < b1:
                b2, b1, l2, l1, bi = b1, d, l1, int(kp_octave[idx]), idx
            elif d < b2:
                l2, b2 = int(kp_octave[idx]), d
        if b1 <= self.TH_HIGH:
            if l1 == l2 and f32(b1) > ratio * f32(b2):
                continue
            assigned[bi] = i
            occupied[bi] = True if has_obs is None else bool(has_obs[i])  # Observations() > 0 of the assigned point (90-92)
            nm += 1
    return nm, assigned


def _search_local_points_device(self, grid, cam, mp_arrays, bad, kp_octave, uright, occupied, th, viewing_cos_limit=0.5):
    """Tracking::SearchLocalPoints (/root/reference/src/Tracking.cc:1156-1206) as one device chain: Frame::isInFrustum for every
    local map point (projection.project, mode 0), the window lookup and all Hamming distances (orbm_window_lists_device) with
    nothing brought to the host in between, then the ordered replay of SearchByProjection(F, vpMapPoints, th)
    (src/ORBmatcher.cc:45-131). mp_arrays: projection.MapPointArrays; bad[i]: pMP->isBad().
    Returns (nmatches, assigned[kp], projection outputs on the host)."""
    from . import projection
    pr, offsets, cands, dist = projection.project_and_window_lists(grid, cam, mp_arrays, 0, viewing_cos_limit, th)
    live = pr["alive"].astype(bool) & ~np.asarray(bad, bool)
    nm, assigned = _replay_frame_mappoints(self, offsets, cands, dist, live, pr["radius"], pr["ur"], kp_octave, uright, occupied)
    return nm, assigned, pr


ORBmatcher.SearchLocalPoints_device = _search_local_points_device
ORBmatcher.SearchByProjection_Frame_MapPoints = _search_by_projection_frame_mappoints
