"""Flat binary form of a corb_ba_problem for the C++ harness tests/host_harness/ba_nccl.cpp (the one-process, N-thread,
N-GPU server path over ncclAllReduce)."""
import numpy as np


def write_problem(path, prob):
    P, L, E = len(prob["pose_fixed"]), len(prob["point_fixed"]), len(prob["edge_pose"])
    with open(path, "wb") as f:
        np.array([P, L, E], np.int32).tofile(f)
        for k, t in (("pose_q", np.float64), ("pose_t", np.float64), ("pose_fixed", np.uint8), ("pose_cam", np.float64),
                     ("point_xyz", np.float64), ("point_fixed", np.uint8), ("edge_pose", np.int32), ("edge_point", np.int32),
                     ("edge_obs", np.float64), ("edge_inv_sigma2", np.float64)):
            np.ascontiguousarray(prob[k], t).tofile(f)
    return P, L, E


def read_result(path, P, L):
    with open(path, "rb") as f:
        n_trials, iterations = np.fromfile(f, np.int32, 2)
        ms, chi2_initial, chi2_final = np.fromfile(f, np.float64, 3)
        acc = np.fromfile(f, np.uint8, min(int(n_trials), 256))
        q = np.fromfile(f, np.float64, 4 * P).reshape(P, 4)
        t = np.fromfile(f, np.float64, 3 * P).reshape(P, 3)
        x = np.fromfile(f, np.float64, 3 * L).reshape(L, 3)
        spread = float(np.fromfile(f, np.float64, 1)[0])
    return {"n_trials": int(n_trials), "iterations": int(iterations), "ms_total": float(ms), "chi2_initial": float(chi2_initial),
            "chi2_final": float(chi2_final), "trial_accepted": [int(v) for v in acc], "pose_q": q, "pose_t": t, "point_xyz": x,
            "rank_spread": spread}
