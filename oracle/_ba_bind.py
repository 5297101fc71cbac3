"""ctypes signatures + numpy wrappers of oracle/ba_oracle.cpp (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C

import numpy as np

_L = None
_f64p = C.POINTER(C.c_double)


class Problem(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("pose_q", C.c_void_p),
                ("pose_t", C.c_void_p), ("pose_fixed", C.c_void_p), ("pose_cam", C.c_void_p), ("point_xyz", C.c_void_p),
                ("point_fixed", C.c_void_p), ("edge_pose", C.c_void_p), ("edge_point", C.c_void_p), ("edge_obs", C.c_void_p),
                ("edge_inv_sigma2", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("n_trials", C.c_int32), ("stopped", C.c_int32), ("solver_failures", C.c_int32),
                ("chi2_initial", C.c_double), ("chi2_final", C.c_double), ("lambda_initial", C.c_double),
                ("lambda_final", C.c_double), ("trial_accepted", C.c_uint8 * 256), ("trial_chi2", C.c_double * 256)]


ALLREDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, _f64p, C.c_int, C.c_int)


def bind(L):
    global _L
    _L = L
    L.oracle_ba_solve.argtypes = [C.POINTER(Problem), C.c_int, C.c_void_p, C.c_int, C.POINTER(Result), C.c_void_p, C.c_void_p]
    L.oracle_ba_edge.argtypes = [_f64p] * 5 + [C.POINTER(C.c_int), _f64p, _f64p, _f64p]
    L.oracle_ba_pose_oplus.argtypes = [_f64p, _f64p, _f64p]
    L.oracle_ba_chi2.restype = C.c_double
    L.oracle_ba_chi2.argtypes = [C.POINTER(Problem), _f64p]


def _lib():
    from . import lib
    lib()
    return _L


_DT = {"pose_q": np.float64, "pose_t": np.float64, "pose_fixed": np.uint8, "pose_cam": np.float64, "point_xyz": np.float64,
       "point_fixed": np.uint8, "edge_pose": np.int32, "edge_point": np.int32, "edge_obs": np.float64,
       "edge_inv_sigma2": np.float64}


def make_problem(prob):
    """dict of arrays (corb_slam_b200.synth.ba_problem layout) -> (ctypes struct, keep-alive dict of contiguous arrays)."""
    keep = {k: np.ascontiguousarray(prob[k], _DT[k]) for k in _DT}
    for k in ("pose_q", "pose_t", "point_xyz"):  # outputs: never write into the caller's arrays
        keep[k] = keep[k].copy()
    p = Problem()
    p.n_poses, p.n_points, p.n_edges = len(keep["pose_fixed"]), len(keep["point_fixed"]), len(keep["edge_pose"])
    for k in _DT:
        setattr(p, k, keep[k].ctypes.data)
    return p, keep


def solve(prob, iterations=10, robust=False, stop=None, allreduce=None):
    """Runs the oracle BA; returns (dict with updated pose_q/pose_t/point_xyz, info dict)."""
    p, keep = make_problem(prob)
    res = Result()
    cb = None
    if allreduce is not None:
        def _cb(user, buf, n, op):
            arr = np.ctypeslib.as_array(buf, shape=(n,)) if n > 0 else np.zeros(0)
            allreduce(arr, op)
            return 0
        cb = ALLREDUCE(_cb)
    stop_p = stop.ctypes.data if stop is not None else None
    rc = _lib().oracle_ba_solve(C.byref(p), iterations, stop_p, int(bool(robust)), C.byref(res),
                                C.cast(cb, C.c_void_p) if cb else None, None)
    assert rc == 0
    n = min(res.n_trials, 256)
    info = {"iterations": res.iterations, "n_trials": res.n_trials, "stopped": res.stopped,
            "solver_failures": res.solver_failures, "chi2_initial": res.chi2_initial, "chi2_final": res.chi2_final,
            "lambda_initial": res.lambda_initial, "lambda_final": res.lambda_final,
            "trial_accepted": [int(res.trial_accepted[i]) for i in range(n)],
            "trial_chi2": [float(res.trial_chi2[i]) for i in range(n)]}
    out = dict(prob)
    out["pose_q"], out["pose_t"], out["point_xyz"] = keep["pose_q"], keep["pose_t"], keep["point_xyz"]
    return out, info


def chi2(prob):
    p, keep = make_problem(prob)
    rms = C.c_double()
    c = _lib().oracle_ba_chi2(C.byref(p), C.byref(rms))
    return c, rms.value


def edge(q, t, cam, X, obs):
    a = [np.ascontiguousarray(v, np.float64) for v in (q, t, cam, X, obs)]
    D = C.c_int()
    e = np.zeros(3); A = np.zeros(9); B = np.zeros(18)
    _lib().oracle_ba_edge(*[v.ctypes.data_as(_f64p) for v in a], C.byref(D), e.ctypes.data_as(_f64p), A.ctypes.data_as(_f64p),
                          B.ctypes.data_as(_f64p))
    d = D.value
    return e[:d].copy(), A[:3 * d].reshape(d, 3).copy(), B[:6 * d].reshape(d, 6).copy()


def pose_oplus(q, t, delta):
    q = np.array(q, np.float64); t = np.array(t, np.float64); d = np.ascontiguousarray(delta, np.float64)
    _lib().oracle_ba_pose_oplus(q.ctypes.data_as(_f64p), t.ctypes.data_as(_f64p), d.ctypes.data_as(_f64p))
    return q, t
