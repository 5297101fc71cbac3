"""Host-side mirror of Optimizer::BundleAdjustment / GlobalBundleAdjustemnt (reference: corbslam_client/include/
Optimizer.h:42-46, src/Optimizer.cc:43-270) over the C ABI. The pointer graph of KeyFrames/MapPoints is passed as the
flat arrays of corb_ba_problem (what the C++ shim builds); results are written back in place.

Multi-GPU: landmarks are sharded over the ranks of a torch.distributed process group, poses are
replicated, and the reduced camera system is all-reduced over NCCL/NVLink from inside corb_ba_solve through a callback."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ALLREDUCE_FN, BaProblem, BaResult, check, lib

_DT = {"pose_q": np.float64, "pose_t": np.float64, "pose_fixed": np.uint8, "pose_cam": np.float64, "point_xyz": np.float64,
       "point_fixed": np.uint8, "edge_pose": np.int32, "edge_point": np.int32, "edge_obs": np.float64,
       "edge_inv_sigma2": np.float64}


class _PageLocked:
    """Owns one corb_host_alloc block; `array` is a numpy view of it."""

    def __init__(self, like):
        like = np.ascontiguousarray(like)
        self._p = C.c_void_p()
        check(lib().corb_host_alloc(max(1, like.nbytes), C.byref(self._p)))
        buf = (C.c_char * max(1, like.nbytes)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=like.dtype, count=like.size).reshape(like.shape)
        self.array[...] = like

    def __del__(self):
        try:
            if self._p:
                lib().corb_host_free(self._p)
                self._p = None
        except Exception:
            pass


def page_locked(problem):
    """A copy of `problem` whose edge arrays live in page-locked memory (corb_host_alloc), like the buffers the C++ shim flattens
    the graph into: corb_ba_solve then uploads them with the DMA engines in place instead of staging them."""
    out = dict(problem)
    keep = []
    for k in ("edge_pose", "edge_point", "edge_obs", "edge_inv_sigma2"):
        h = _PageLocked(np.asarray(problem[k], _DT[k]))
        keep.append(h)
        out[k] = h.array
    out["_page_locked"] = keep  # owners of the blocks: live as long as the problem dict
    return out


class _DevArray:
    """Wraps a raw device pointer as a __cuda_array_interface__ object so torch can view it without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def torch_allreduce(group=None, device=None):
    """All-reduce callback backed by torch.distributed (NCCL over NVLink on the GPU box). `device`: the CUDA device index
    corb_ba_solve runs on (default: torch's current device at the time this is called)."""
    import torch
    import torch.distributed as dist
    ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))

    def _cb(user, d_buf, n, op, stream):
        try:
            with torch.cuda.device(dev):
                t = torch.as_tensor(_DevArray(d_buf, n), device=dev)
                if t.data_ptr() != d_buf:  # a silent copy would leave the solver's own buffer unreduced
                    raise RuntimeError("the all-reduce buffer was copied (device mismatch)")
                s = torch.cuda.ExternalStream(stream, device=dev)
                with torch.cuda.stream(s):
                    dist.all_reduce(t, op=ops[op], group=group)
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            print("corb all-reduce callback failed:", repr(e))
            return 1
    return ALLREDUCE_FN(_cb)


def release_cache(device=0):
    """Returns the device memory corb_ba_solve keeps between calls (corb_ba_release_cache)."""
    check(lib().corb_ba_release_cache(int(device)))


class Optimizer:
    @staticmethod
    def BundleAdjustment(problem, nIterations=5, pbStopFlag=None, nLoopKF=0, bRobust=True, device=0, allreduce=None):
        """problem: dict with the arrays of corb_ba_problem (see corb_slam_b200.synth.ba_problem). The returned dict
        holds updated copies of pose_q / pose_t / point_xyz (the reference writes mTcwGBA / mPosGBA, Optimizer.cc:219-263). pbStopFlag: optional np.uint8 array of one element polled by the solver.
        Returns (problem_out, info)."""
        keep = {k: np.ascontiguousarray(problem[k], _DT[k]) for k in _DT}
        for k in ("pose_q", "pose_t", "point_xyz"):  # outputs are returned as copies, inputs stay untouched
            keep[k] = keep[k].copy()
        p = BaProblem()
        p.n_poses, p.n_points, p.n_edges = len(keep["pose_fixed"]), len(keep["point_fixed"]), len(keep["edge_pose"])
        for k in _DT:
            setattr(p, k, keep[k].ctypes.data)
        res = BaResult()
        stop_p = pbStopFlag.ctypes.data if pbStopFlag is not None else None
        cb = C.cast(allreduce, C.c_void_p) if allreduce is not None else None
        st = lib().corb_ba_solve(C.byref(p), int(nIterations), stop_p, int(bool(bRobust)), int(device), C.byref(res), cb, None)
        if st not in (_lib.OK, _lib.ERR_STOPPED):
            check(st)
        n = min(res.n_trials, 256)
        info = {"iterations": res.iterations, "n_trials": res.n_trials, "stopped": res.stopped,
                "solver_failures": res.solver_failures, "chi2_initial": res.chi2_initial, "chi2_final": res.chi2_final,
                "lambda_initial": res.lambda_initial, "lambda_final": res.lambda_final,
                "trial_accepted": [int(res.trial_accepted[i]) for i in range(n)],
                "trial_chi2": [float(res.trial_chi2[i]) for i in range(n)], "ms_total": res.ms_total,
                "ms_solve": res.ms_solve, "reduced_blocks": res.reduced_blocks, "border_poses": res.border_poses,
                "max_active_rows": res.max_active_rows, "ms_setup": res.ms_setup, "band_chunks": res.band_chunks,
                "separator_poses": res.separator_poses, "schur_pair_lists": res.schur_pair_lists}
        out = dict(problem)
        out["pose_q"], out["pose_t"], out["point_xyz"] = keep["pose_q"], keep["pose_t"], keep["point_xyz"]
        return out, info

    @staticmethod
    def GlobalBundleAdjustemnt(problem, nIterations=5, pbStopFlag=None, nLoopKF=0, bRobust=True, device=0, allreduce=None):
        """Same spelling as the reference (Optimizer.cc:43)."""
        return Optimizer.BundleAdjustment(problem, nIterations, pbStopFlag, nLoopKF, bRobust, device, allreduce)
