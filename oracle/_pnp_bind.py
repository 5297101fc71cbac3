"""ctypes signatures + numpy wrappers of oracle/pnp_oracle.cpp (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C

import numpy as np

_L = None
_f64p = C.POINTER(C.c_double)


def bind(L):
    global _L
    _L = L
    vp = C.c_void_p
    L.oracle_pnp_ransac_params.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp]
    L.oracle_pnp_ransac_params.restype = None
    L.oracle_pnp_create.argtypes = [C.c_int, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
    L.oracle_pnp_create.restype = vp
    L.oracle_pnp_destroy.argtypes = [vp]
    L.oracle_pnp_destroy.restype = None
    L.oracle_pnp_iterations.argtypes = [vp]
    L.oracle_pnp_iterate.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp]
    L.oracle_epnp_pose.argtypes = [C.c_int, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp]
    L.oracle_epnp_pose.restype = C.c_double
    L.oracle_svd.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.oracle_svd.restype = None
    L.oracle_svd_solve.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.oracle_svd_solve.restype = None
    L.oracle_svd_invert3.argtypes = [vp, vp]
    L.oracle_svd_invert3.restype = None


def _lib():
    from . import lib
    lib()
    return _L


def ransac_params(N, probability=0.99, min_inliers=8, max_iterations=300, min_set=4, epsilon=0.4):
    """PnPsolver::SetRansacParameters -> (adjusted mRansacMinInliers, mRansacMaxIts)."""
    a, b = C.c_int(0), C.c_int(0)
    _lib().oracle_pnp_ransac_params(int(N), float(probability), int(min_inliers), int(max_iterations), int(min_set), float(epsilon),
                                    C.addressof(a), C.addressof(b))
    return a.value, b.value


class PnpSolver:
    """Sequential restatement of PnPsolver (state kept between iterate() calls). `draws[it, k]` is what RandomInt returned."""

    def __init__(self, p2d, p3d, max_err, fx, fy, cx, cy, min_inliers, max_its):
        self.p2d = np.ascontiguousarray(p2d, np.float32).reshape(-1, 2)
        self.p3d = np.ascontiguousarray(p3d, np.float32).reshape(-1, 3)
        self.max_err = np.ascontiguousarray(max_err, np.float32)
        self.N = len(self.p2d)
        self.h = _lib().oracle_pnp_create(self.N, self.p2d.ctypes.data, self.p3d.ctypes.data, self.max_err.ctypes.data, fx, fy, cx, cy,
                                          int(min_inliers), int(max_its))

    def iterate(self, n_iterations, draws):
        draws = np.ascontiguousarray(draws, np.int32).reshape(-1, 4)
        no_more, n_inl = C.c_int(0), C.c_int(0)
        inl = np.zeros(max(self.N, 1), np.uint8)
        T = np.zeros(16, np.float32)
        rc = _lib().oracle_pnp_iterate(self.h, int(n_iterations), draws.ctypes.data, len(draws), C.addressof(no_more), inl.ctypes.data,
                                       C.addressof(n_inl), T.ctypes.data)
        if rc < 0:
            raise ValueError("not enough draws")
        return rc, bool(no_more.value), inl[:self.N].astype(bool), n_inl.value, T.reshape(4, 4)

    @property
    def iterations(self):
        return _lib().oracle_pnp_iterations(self.h)

    def __del__(self):
        if getattr(self, "h", None) and _L is not None:
            _L.oracle_pnp_destroy(self.h)
            self.h = None


def epnp_pose(pws, us, fu, fv, uc, vc):
    pws = np.ascontiguousarray(pws, np.float64).reshape(-1, 3)
    us = np.ascontiguousarray(us, np.float64).reshape(-1, 2)
    R, t = np.zeros(9), np.zeros(3)
    err = _lib().oracle_epnp_pose(len(pws), pws.ctypes.data, us.ctypes.data, fu, fv, uc, vc, R.ctypes.data, t.ctypes.data)
    return R.reshape(3, 3), t, err


def svd(A):
    """(Ut, W, Vt) of a row-major m x n matrix, m >= n: rows of Ut / Vt are the left / right singular vectors."""
    A = np.ascontiguousarray(A, np.float64)
    m, n = A.shape
    Ut, W, Vt = np.zeros((n, m)), np.zeros(n), np.zeros((n, n))
    _lib().oracle_svd(A.ctypes.data, m, n, Ut.ctypes.data, W.ctypes.data, Vt.ctypes.data)
    return Ut, W, Vt


def svd_solve(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(A.shape[1])
    _lib().oracle_svd_solve(A.ctypes.data, A.shape[0], A.shape[1], b.ctypes.data, x.ctypes.data)
    return x


def svd_invert3(A):
    A = np.ascontiguousarray(A, np.float64)
    inv = np.zeros((3, 3))
    _lib().oracle_svd_invert3(A.ctypes.data, inv.ctypes.data)
    return inv
