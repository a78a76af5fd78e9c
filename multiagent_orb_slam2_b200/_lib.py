"""ctypes binding of the C ABI (include/orb_b200.h). There is no CPU fallback: a missing or
unloadable liborb_b200.so is a hard error, and every compute call fails without a CUDA device."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liborb_b200.so")

ORB_OK, ORB_EINVAL, ORB_ECUDA, ORB_ECAPACITY = 0, -1, -2, -3


class OrbError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("orb_b200 error %d: %s" % (code, text))
        self.code = code


class Config(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrbError(ORB_ECUDA, "CUDA extension %s is missing - run `python -m multiagent_orb_slam2_b200.build` "
                                  "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, sz, f32 = C.c_void_p, C.c_int, C.c_size_t, C.c_float
    L.orb_last_error.restype = C.c_char_p
    L.orb_version.restype = C.c_char_p
    sigs = {
        "orb_device_count": [],
        "orbx_create": [C.POINTER(Config), i32, i32, i32, i32, C.POINTER(vp)],
        "orbx_get_tables": [vp, vp, vp, vp, vp, vp],
        "orbx_max_keypoints": [vp],
        "orbx_set_stage_timing": [vp, i32],
        "orbx_stage_times": [vp, vp],
        "orbx_algorithmic_bytes": [vp, vp],
        "orbx_level_size": [vp, i32, C.POINTER(i32), C.POINTER(i32)],
        "orbx_extract": [vp, vp, sz, vp, vp, i32, C.POINTER(i32)],
        "orbx_extract_batch": [vp, vp, sz, sz, i32, vp, vp, i32, vp],
        "orbx_extract_device": [vp, vp, sz, sz, i32, vp],
        "orbx_upload_frames": [vp, vp, sz, sz, i32, vp],
        "orbx_extract_staged": [vp, i32, vp],
        "orbx_download_results": [vp, i32, vp, vp, i32, vp, vp],
        "orbx_device_results": [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32)],
        "orbx_pyramid_level_device": [vp, i32, i32, C.POINTER(vp), C.POINTER(sz)],
        "orbx_pyramid_level": [vp, i32, i32, vp, sz],
        "orbx_pyramid_levels": [vp, i32, i32, i32, vp, vp],
        "orbx_debug_blurred_level": [vp, i32, i32, vp, sz],
        "orbx_debug_candidates": [vp, i32, i32, vp, i32, C.POINTER(i32)],
        "orbx_debug_quadtree": [i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, C.POINTER(i32)],
        "orbm_knn2_device": [vp, i32, vp, i32, vp, vp, vp, vp],
        "orbm_knn2": [i32, vp, i32, vp, i32, vp, vp, vp],
        "orbm_release_scratch": [vp],
        "orbm_knn2_batched_device": [vp, vp, i32, vp, vp, i32, i32, vp, vp, vp, vp],
        "orbm_knn2_pairs_device": [vp, vp, i32, vp, i32, vp, vp, vp, vp],
        "orbm_knn2_lists_device": [vp, i32, vp, vp, vp, vp, vp, vp, vp],
        "orbm_knn2_lists": [i32, vp, i32, vp, i32, vp, vp, vp, vp, vp],
        "orbm_list_distances_device": [vp, i32, vp, vp, vp, vp, vp],
        "orbm_list_distances": [i32, vp, i32, vp, i32, vp, vp, vp],
        "orbm_ratio_filter_device": [vp, vp, vp, i32, i32, i32, f32, vp, vp],
        "orbm_stereo_match_device": [vp, vp, i32, f32, f32, vp, vp, vp, vp, vp],
        "orbm_stereo_match_batch_device": [vp, vp, i32, f32, f32, vp, vp, vp, vp, i32, vp],
        "orbm_stereo_match": [vp, vp, i32, f32, f32, vp, vp, i32, C.POINTER(i32)],
        "orbm_grid_create": [i32, i32, C.POINTER(vp)],
        "orbm_grid_build_device": [vp, vp, vp, f32, f32, f32, f32, vp],
        "orbm_window_knn2_device": [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "orbm_window_lists_device": [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(i32), vp],
        "orbm_window_lists": [i32, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(i32)],
        "orbv_create": [i32, vp, vp, vp, i32, i32, i32, C.POINTER(vp)],
        "orbv_descend_device": [vp, vp, i32, i32, vp, vp, vp, vp],
        "orbv_descend": [vp, vp, i32, i32, vp, vp, vp],
        "orbm_distance_matrix_device": [vp, i32, vp, i32, vp, vp],
        "orbm_distance_matrix": [i32, vp, i32, vp, i32, vp],
        "orbm_set_knn2_backend": [i32],
        "orbm_debug_mma_occupancy": [C.POINTER(i32), C.POINTER(i32)],
        "orbm_knn2_mma_device": [vp, i32, vp, i32, vp, vp, vp, vp],
        "orbm_debug_mma_dot": [i32, vp, vp, vp],
        "orbdb_create": [i32, i32, C.POINTER(vp)],
        "orbdb_add": [vp, vp, vp, i32, C.POINTER(i32)],
        "orbdb_erase": [vp, i32],
        "orbdb_size": [vp],
        "orbdb_query": [vp, vp, vp, i32, vp, vp, vp, i32],
        "orbdb_query_device": [vp, vp, vp, i32, vp, vp, vp, vp],
        "orbm_stereo_from_rgbd_device": [i32, vp, vp, vp, i32, vp, i32, i32, C.c_size_t, f32, vp, vp, vp],
        "orbm_undistort_keypoints_device": [i32, vp, vp, i32, f32, f32, f32, f32, vp, i32, vp, vp],
        "orbm_image_bounds": [i32, i32, f32, f32, f32, f32, vp, i32, vp],
        "orbm_xmap_create": [i32, i32, i32, i32, i32, C.POINTER(vp)],
        "orbm_xmap_ipc_handle": [vp, vp],
        "orbm_xmap_attach_ipc": [vp, vp],
        "orbm_xmap_attach_local": [vp, i32],
        "orbm_knn2_allgather": [vp, vp, vp, vp],
        "orbm_xmap_result": [vp, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
        "orbm_xmap_plan": [i32, vp, i32, i32, vp, i32, C.POINTER(i32)],
        "orbm_project_points_device": [i32, vp, i32, f32, f32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.orb_launch_count.argtypes = []
    L.orb_launch_count.restype = C.c_longlong
    L.orbm_grid_destroy.argtypes = [vp]
    L.orbm_grid_destroy.restype = None
    L.orbdb_destroy.argtypes = [vp]
    L.orbdb_destroy.restype = None
    L.orbv_destroy.argtypes = [vp]
    L.orbv_destroy.restype = None
    L.orbx_destroy.argtypes = [vp]
    L.orbx_destroy.restype = None
    L.orbm_xmap_destroy.argtypes = [vp]
    L.orbm_xmap_destroy.restype = None
    _lib = L
    return L


def check(rc, allow=()):
    if rc != ORB_OK and rc not in allow:
        raise OrbError(rc, lib().orb_last_error().decode())
    return rc


def exported_symbols():
    """Names declared in include/orb_b200.h and include/orb_b200_debug.h (parsed), for the load/export test."""
    import re
    names = set()
    for h in ("orb_b200.h", "orb_b200_debug.h"):
        hdr = open(os.path.join(HERE, "..", "include", h)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        names |= set(re.findall(r"\b(orb[a-z]*_[a-z0-9_]+)\s*\(", hdr))
    return sorted(names)
