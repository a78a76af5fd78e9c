"""Device-side map-point projection: Frame::isInFrustum (/root/reference/src/Frame.cc:269-325) and
the projection prologue of the Fuse / Sim3 searches for a structure-of-arrays mirror of MapPoint::mWorldPos /
mNormalVector / mfMaxDistance / mfMinDistance, chained into the windowed search of DeviceFrameGrid without a host round
trip. torch is plumbing (device arrays)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

MAX_LEVELS = 32


class Camera(C.Structure):
    """orbm_camera of include/orb_b200.h."""
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("min_x", C.c_float), ("max_x", C.c_float),
                ("min_y", C.c_float), ("max_y", C.c_float), ("log_scale_factor", C.c_float), ("n_levels", C.c_int32),
                ("scale_factors", C.c_float * MAX_LEVELS)]

    @classmethod
    def make(cls, Rcw, tcw, Ow, fx, fy, cx, cy, bf, bounds, scale_factors, log_scale_factor=None):
        c = cls()
        c.Rcw[:] = [float(x) for x in np.asarray(Rcw, np.float32).reshape(9)]
        c.tcw[:] = [float(x) for x in np.asarray(tcw, np.float32).reshape(3)]
        c.Ow[:] = [float(x) for x in np.asarray(Ow, np.float32).reshape(3)]
        c.fx, c.fy, c.cx, c.cy, c.bf = float(fx), float(fy), float(cx), float(cy), float(bf)
        c.min_x, c.max_x, c.min_y, c.max_y = [float(b) for b in bounds]
        sf = np.asarray(scale_factors, np.float32)
        c.n_levels = len(sf)
        for i, s in enumerate(sf):
            c.scale_factors[i] = float(s)
        # mfLogScaleFactor = log(mfScaleFactor) (src/Frame.cc:71): std::log(float), a float
        c.log_scale_factor = float(np.log(sf[1]).astype(np.float32)) if log_scale_factor is None else float(log_scale_factor)
        return c


class MapPointArrays:
    """SoA mirror of the map points a search projects: world position, normal (n x 3 float32), mfMaxDistance, mfMinDistance
    and the distinctive descriptor (n x 32 uint8), resident on the device."""

    def __init__(self, device, world_pos, normal, max_distance, min_distance, desc):
        self.dev = torch.device("cuda", device)
        up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dt)).to(self.dev)
        self.n = len(world_pos)
        self.pos, self.normal = up(world_pos, np.float32), up(normal, np.float32)
        self.max_d, self.min_d = up(max_distance, np.float32), up(min_distance, np.float32)
        self.desc = up(desc, np.uint8)


def project(device, cam, mp, mode=0, viewing_cos_limit=0.5, th=1.0):
    """orbm_project_points_device on a MapPointArrays; returns a dict of device tensors alive, u, v, ur, level, view_cos,
    radius, min_level, max_level (mode 0 = isInFrustum, mode 1 = the Fuse / Sim3 prologue)."""
    L = _lib.lib()
    dev = torch.device("cuda", device)
    n = mp.n
    out = dict(alive=torch.empty(n, dtype=torch.uint8, device=dev))
    for k in ("u", "v", "ur", "view_cos", "radius"):
        out[k] = torch.empty(n, dtype=torch.float32, device=dev)
    for k in ("level", "min_level", "max_level"):
        out[k] = torch.empty(n, dtype=torch.int32, device=dev)
    p = lambda z: C.c_void_p(z.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(L.orbm_project_points_device(device, C.byref(cam), mode, float(viewing_cos_limit), float(th), p(mp.pos), p(mp.normal),
                                            p(mp.max_d), p(mp.min_d), n, p(out["alive"]), p(out["u"]), p(out["v"]), p(out["ur"]),
                                            p(out["level"]), p(out["view_cos"]), p(out["radius"]), p(out["min_level"]),
                                            p(out["max_level"]), st))
    return out


def project_and_window_lists(grid, cam, mp, mode=0, viewing_cos_limit=0.5, th=1.0):
    """Projection -> window lookup -> Hamming on the device in one chain (no host data in between): returns the host
    copies of the projection outputs and the CSR candidate lists (offsets, cands, dist) for the ordered replay."""
    L = _lib.lib()
    pr = project(grid.ex.device, cam, mp, mode, viewing_cos_limit, th)
    n = mp.n
    dev = pr["u"].device
    p = lambda z: C.c_void_p(z.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    cap = 64 * max(n, 1)
    while True:
        offs = torch.empty(n + 1, dtype=torch.int32, device=dev)
        cands = torch.empty(cap, dtype=torch.int32, device=dev)
        dist = torch.empty(cap, dtype=torch.int16, device=dev)
        total = C.c_int()
        rc = L.orbm_window_lists_device(grid._g, C.c_void_p(grid.d_desc), p(mp.desc), n, p(pr["u"]), p(pr["v"]), p(pr["radius"]),
                                        p(pr["min_level"]), p(pr["max_level"]), p(offs), p(cands), p(dist), cap, C.byref(total), st)
        if rc == _lib.ORB_ECAPACITY:
            cap = total.value
            continue
        _lib.check(rc)
        break
    torch.cuda.synchronize(dev)
    host = {k: v.cpu().numpy() for k, v in pr.items()}
    return host, offs.cpu().numpy(), cands[:total.value].cpu().numpy(), dist[:total.value].cpu().numpy()
