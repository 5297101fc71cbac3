"""ctypes bindings of oracle/_ref/libshim.so (TEST INFRASTRUCTURE ONLY): the reference's own Frame.cc, the rest of ORBmatcher.cc
and PnPsolver.cc, DBoW2 - with the hot-path bodies taken over by shim/*.cc and linked against corb_slam_b200/libcorb_b200.so.
Same glue, same Python classes as oracle/ref.py (that module's source is instantiated a second time on the other library), so
a test drives the reference CPU code and the GPU drop-in through the SAME C++ class seams, plus the Optimizer::BundleAdjustment
seam of shim/Optimizer_gba.cc."""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libshim.so")


def available():
    return os.path.exists(LIB_PATH)


def _instantiate():
    spec = importlib.util.spec_from_file_location("oracle._ref_on_shim", os.path.join(_HERE, "ref.py"))
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "oracle"
    sys.modules["oracle._ref_on_shim"] = mod
    spec.loader.exec_module(mod)
    mod.LIB_PATH = LIB_PATH
    return mod


classes = _instantiate()  # classes.ORBextractor, classes.Frame, classes.search_by_bow, ... exactly as in oracle.ref
_vp, _i32p = C.c_void_p, C.POINTER(C.c_int32)
_bound = False


def lib():
    global _bound
    L = classes.lib()
    if not _bound:
        L.shim_world_create.restype = _vp
        L.shim_world_create.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int]
        L.shim_world_destroy.argtypes = [_vp]
        L.shim_ba_flatten.argtypes = [_vp, _i32p, _i32p, _i32p]
        L.shim_ba_flat_get.argtypes = [_vp] * 13
        L.shim_ba_flat_set_and_write_back.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64]
        L.shim_ba_run.argtypes = [_vp, C.c_int, C.c_uint64, C.c_int]
        L.shim_world_read.argtypes = [_vp] * 9
        L.shim_quat_from_pose.argtypes = [_vp, _vp, _vp]
        L.shim_pose_from_quat.argtypes = [_vp, _vp, _vp]
        _bound = True
    return L


def quat_from_pose(T):
    T = np.ascontiguousarray(T, np.float32).reshape(16)
    q, t = np.zeros(4), np.zeros(3)
    lib().shim_quat_from_pose(T.ctypes.data, q.ctypes.data, t.ctypes.data)
    return q, t


def pose_from_quat(q, t):
    q = np.ascontiguousarray(q, np.float64); t = np.ascontiguousarray(t, np.float64)
    T = np.zeros(16, np.float32)
    lib().shim_pose_from_quat(q.ctypes.data, t.ctypes.data, T.ctypes.data)
    return T.reshape(4, 4)


class World:
    """KeyFrame / MapPoint objects (stand-ins) built from flat arrays, the inputs of Optimizer::BundleAdjustment."""

    def __init__(self, kf_id, kf_Tcw, kf_flags, kf_cam, mp_xyz, mp_flags, obs_kf, obs_mp, obs_uvr, obs_octave, inv_sigma2):
        a = lambda v, t: np.ascontiguousarray(v, t)
        self.kf_id, self.kf_Tcw, self.kf_flags = a(kf_id, np.uint64), a(kf_Tcw, np.float32).reshape(-1, 16), a(kf_flags, np.uint8)
        self.kf_cam, self.mp_xyz, self.mp_flags = a(kf_cam, np.float32).reshape(-1, 5), a(mp_xyz, np.float32).reshape(-1, 3), a(mp_flags, np.uint8)
        self.obs_kf, self.obs_mp = a(obs_kf, np.int32), a(obs_mp, np.int32)
        self.obs_uvr, self.obs_octave, self.inv_sigma2 = a(obs_uvr, np.float32).reshape(-1, 3), a(obs_octave, np.int32), a(inv_sigma2, np.float32)
        self.n_kf, self.n_mp = len(self.kf_id), len(self.mp_xyz)
        p = lambda v: v.ctypes.data
        self._h = lib().shim_world_create(self.n_kf, p(self.kf_id), p(self.kf_Tcw), p(self.kf_flags), p(self.kf_cam), self.n_mp, p(self.mp_xyz),
                                          p(self.mp_flags), len(self.obs_kf), p(self.obs_kf), p(self.obs_mp), p(self.obs_uvr),
                                          p(self.obs_octave), p(self.inv_sigma2), len(self.inv_sigma2))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().shim_world_destroy(self._h)
            self._h = None

    def flatten(self):
        """FlatBA::build (Optimizer.cc:80-207) -> dict with the arrays of corb_ba_problem + the dense index maps."""
        nP, nL, nE = C.c_int32(), C.c_int32(), C.c_int32()
        lib().shim_ba_flatten(self._h, C.byref(nP), C.byref(nL), C.byref(nE))
        P, L, E = nP.value, nL.value, nE.value
        z = lambda n, t: np.zeros(max(n, 1), t)
        f = {"pose_q": z(4 * P, np.float64), "pose_t": z(3 * P, np.float64), "pose_fixed": z(P, np.uint8), "pose_cam": z(5 * P, np.float64),
             "point_xyz": z(3 * L, np.float64), "point_fixed": z(L, np.uint8), "edge_pose": z(E, np.int32), "edge_point": z(E, np.int32),
             "edge_obs": z(3 * E, np.float64), "edge_inv_sigma2": z(E, np.float64), "pose_kf_id": z(P, np.uint64), "point_mp_index": z(L, np.int32)}
        order = ["pose_q", "pose_t", "pose_fixed", "pose_cam", "point_xyz", "point_fixed", "edge_pose", "edge_point", "edge_obs",
                 "edge_inv_sigma2", "pose_kf_id", "point_mp_index"]
        lib().shim_ba_flat_get(self._h, *[f[k].ctypes.data for k in order])
        n = {"pose_q": (P, 4), "pose_t": (P, 3), "pose_fixed": (P,), "pose_cam": (P, 5), "point_xyz": (L, 3), "point_fixed": (L,),
             "edge_pose": (E,), "edge_point": (E,), "edge_obs": (E, 3), "edge_inv_sigma2": (E,), "pose_kf_id": (P,), "point_mp_index": (L,)}
        return {k: f[k][:int(np.prod(n[k]))].reshape(n[k]).copy() for k in order}

    def write_back(self, pose_q, pose_t, point_xyz, nLoopKF):
        a = lambda v: np.ascontiguousarray(v, np.float64)
        q, t, x = a(pose_q), a(pose_t), a(point_xyz)
        lib().shim_ba_flat_set_and_write_back(self._h, q.ctypes.data, t.ctypes.data, x.ctypes.data, int(nLoopKF))

    def run(self, nIterations, nLoopKF, bRobust):
        """Optimizer::BundleAdjustment(vpKFs, vpMP, nIterations, NULL, nLoopKF, bRobust) through the shim, on the GPU."""
        lib().shim_ba_run(self._h, int(nIterations), int(nLoopKF), int(bool(bRobust)))

    def read(self):
        T, G = np.zeros((self.n_kf, 16), np.float32), np.zeros((self.n_kf, 16), np.float32)
        kg = np.zeros(self.n_kf, np.uint64)
        X, XG = np.zeros((self.n_mp, 3), np.float32), np.zeros((self.n_mp, 3), np.float32)
        mg, nu, cc = np.zeros(self.n_mp, np.uint64), np.zeros(self.n_mp, np.int32), np.zeros(2, np.int32)
        lib().shim_world_read(self._h, *[v.ctypes.data for v in (T, G, kg, X, XG, mg, nu, cc)])
        return {"Tcw": T.reshape(-1, 4, 4), "TcwGBA": G.reshape(-1, 4, 4), "kf_gba": kg, "xyz": X, "posGBA": XG, "mp_gba": mg,
                "normal_updates": nu, "cache_kfs": int(cc[0]), "cache_mps": int(cc[1])}
