"""ctypes loader of libcorb_b200.so (the C ABI declared in include/corb_b200.h).

There is no fallback: if the shared library is missing the import fails, and every compute entry point fails
with CORB_ERR_CUDA when no B200 is visible.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CORB_LIB selects another BUILD of the same library (instrumented debug variants, csrc/Makefile); there is no other backend
LIB_PATH = os.environ.get("CORB_LIB") or os.path.join(_HERE, "libcorb_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_CAPACITY, ERR_STOPPED, ERR_IO = range(7)

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
vp = C.c_void_p


class BowSide(C.Structure):
    """corb_bow_side (include/corb_b200.h)."""
    _fields_ = [("desc", vp), ("n", C.c_int32), ("fv_nodes", vp), ("fv_off", vp), ("fv_idx", vp), ("fv_n", C.c_int32),
                ("valid", vp), ("angles", vp)]


class BaProblem(C.Structure):
    """corb_ba_problem (include/corb_b200.h)."""
    _fields_ = [("n_poses", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("pose_q", vp), ("pose_t", vp),
                ("pose_fixed", vp), ("pose_cam", vp), ("point_xyz", vp), ("point_fixed", vp), ("edge_pose", vp),
                ("edge_point", vp), ("edge_obs", vp), ("edge_inv_sigma2", vp)]


class BaResult(C.Structure):
    """corb_ba_result (include/corb_b200.h)."""
    _fields_ = [("iterations", C.c_int32), ("n_trials", C.c_int32), ("stopped", C.c_int32), ("solver_failures", C.c_int32),
                ("chi2_initial", C.c_double), ("chi2_final", C.c_double), ("lambda_initial", C.c_double),
                ("lambda_final", C.c_double), ("trial_accepted", C.c_uint8 * 256), ("trial_chi2", C.c_double * 256),
                ("ms_total", C.c_double), ("ms_solve", C.c_double), ("reduced_blocks", C.c_int64),
                ("border_poses", C.c_int32), ("max_active_rows", C.c_int32), ("ms_setup", C.c_double),
                ("band_chunks", C.c_int32), ("separator_poses", C.c_int32), ("schur_pair_lists", C.c_int32), ("reserved0", C.c_int32)]


class PnpProblem(C.Structure):
    """corb_pnp_problem (include/corb_b200.h)."""
    _fields_ = [("n", C.c_int32), ("p2d", vp), ("p3d", vp), ("max_err", vp), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("min_inliers", C.c_int32), ("max_its", C.c_int32), ("iterations_done", C.c_int32),
                ("n_iterations", C.c_int32), ("draws", vp)]


class PnpResult(C.Structure):
    """corb_pnp_result (include/corb_b200.h)."""
    _fields_ = [("status", C.c_int32), ("no_more", C.c_int32), ("n_inliers", C.c_int32), ("iterations", C.c_int32),
                ("Tcw", C.c_float * 16)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, vp, vp, C.c_size_t, C.c_int, vp)


class CorbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("corb_b200 error %d: %s" % (status, msg))
        self.status = status


def build(verbose=False):
    """Compile libcorb_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libcorb_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing - run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.corb_last_error.restype = C.c_char_p
        L.corb_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.corb_orb_destroy.argtypes = [vp]
        L.corb_orb_destroy.restype = None
        L.corb_orb_levels.argtypes = [vp]
        L.corb_orb_scale_factor.argtypes = [vp]
        L.corb_orb_scale_factor.restype = C.c_float
        L.corb_orb_tables.argtypes = [vp, f32p, f32p, f32p, f32p, i32p, i32p]
        L.corb_orb_level_size.argtypes = [vp, C.c_int, C.c_int, C.c_int, i32p, i32p]
        L.corb_orb_capacity.argtypes = [vp, C.c_int, C.c_int]
        L.corb_orb_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, i32p, C.POINTER(vp)]
        L.corb_orb_extract_submit.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.corb_orb_extract_wait.argtypes = [vp, vp, vp, i32p, C.POINTER(vp)]
        L.corb_orb_extract_pair.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, i32p, vp, vp, i32p, C.POINTER(vp),
                                            C.POINTER(vp)]
        L.corb_orb_extract_pair_device.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int]
        L.corb_stereo_match.argtypes = [vp, vp, C.c_float, C.c_float, C.c_int, vp, vp]
        L.corb_frame_stereo.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp, i32p, vp, vp, i32p, vp, vp]
        L.corb_bow_store_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.corb_bow_store_destroy.argtypes = [vp]
        L.corb_bow_store_destroy.restype = None
        L.corb_bow_store_fill.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp]
        L.corb_frame_bow.argtypes = [vp, vp, C.c_int, vp]
        L.corb_bow_store_side.argtypes = [vp, vp, C.POINTER(BowSide), i32p]
        L.corb_bow_store_sync.argtypes = [vp]
        L.corb_bow_store_features.argtypes = [vp]
        L.corb_bow_store_download.argtypes = [vp, vp, vp, i32p, vp, vp, vp, i32p]
        L.corb_bow_score_stores.argtypes = [vp, vp, C.c_int, C.POINTER(vp), vp]
        L.corb_bow_match_stores.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.c_float, C.c_int,
                                            C.POINTER(vp), i32p]
        L.corb_orb_extract_pair_submit.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.corb_orb_extract_pair_wait.argtypes = [vp, vp, vp, vp, i32p, vp, vp, i32p, C.POINTER(vp), C.POINTER(vp)]
        L.corb_frame_stereo_submit.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        L.corb_frame_stereo_wait.argtypes = [vp, vp, vp, vp, i32p, vp, vp, i32p, vp, vp]
        L.corb_orb_extract_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.corb_orb_sync.argtypes = [vp]
        L.corb_orb_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.corb_orb_host_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), i32p]
        L.corb_orb_device_level.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), i32p, i32p, i32p]
        L.corb_orb_stream.argtypes = [vp]
        L.corb_orb_stream.restype = vp
        L.corb_orb_launches_per_extract.argtypes = [vp]
        L.corb_orb_uses_tma.argtypes = [vp]
        L.corb_orb_set_host_transfer.argtypes = [vp, C.c_int]
        L.corb_orb_host_transfer.argtypes = [vp]
        L.corb_orb_profile.argtypes = [vp, C.c_int, f32p, C.c_int, i32p]
        L.corb_orb_kernel_name.argtypes = [vp, C.c_int]
        L.corb_orb_kernel_name.restype = C.c_char_p
        L.corb_orb_tap.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t, i32p]
        u32p = C.POINTER(C.c_uint32)
        L.corb_matcher_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.corb_matcher_destroy.argtypes = [vp]
        L.corb_matcher_destroy.restype = None
        L.corb_matcher_sync.argtypes = [vp]
        L.corb_matcher_stream.argtypes = [vp]
        L.corb_matcher_stream.restype = vp
        L.corb_hamming_pairs.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]
        L.corb_bow_match.argtypes = [vp, C.c_int, C.POINTER(BowSide), C.POINTER(BowSide), C.c_float, C.c_int, vp, i32p]
        L.corb_bow_match_batch.argtypes = [vp, C.c_int, C.c_int, C.POINTER(BowSide), C.POINTER(BowSide), C.c_float, C.c_int,
                                           C.POINTER(vp), i32p]
        L.corb_bow_match_batch_device.argtypes = [vp, C.c_int, C.c_int, C.POINTER(BowSide), C.POINTER(BowSide), C.c_float,
                                                  C.c_int, C.POINTER(vp), vp]
        L.corb_search_by_projection_last.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp, i32p]
        L.corb_search_by_projection_map.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, i32p]
        L.corb_voc_load_text.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.corb_voc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.POINTER(vp)]
        L.corb_voc_destroy.argtypes = [vp]
        L.corb_voc_destroy.restype = None
        L.corb_voc_info.argtypes = [vp] + [i32p] * 6
        L.corb_voc_transform_features.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
        L.corb_voc_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, i32p, vp, vp, vp, i32p]
        L.corb_bow_score_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), vp, vp]
        L.corb_pnp_ransac_params.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_float, i32p, i32p]
        L.corb_pnp_iterate_batch.argtypes = [vp, C.c_int, C.POINTER(PnpProblem), C.POINTER(PnpResult), C.POINTER(vp)]
        L.corb_ba_release_cache.argtypes = [C.c_int]
        L.corb_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.corb_host_free.argtypes = [C.c_void_p]
        L.corb_ba_solve.argtypes = [C.POINTER(BaProblem), C.c_int, vp, C.c_int, C.c_int, C.POINTER(BaResult), vp, vp]
        _lib = L
    return _lib


def check(status):
    if status != OK:
        raise CorbError(status, lib().corb_last_error().decode("utf-8", "replace"))
