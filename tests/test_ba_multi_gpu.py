"""2-GPU run of the landmark-sharded BA with the NCCL all-reduce hook (skipped on boxes with one GPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_ba_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "ba_multi_check.py"), "100", "8000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _harness():
    exe = os.path.join(ROOT, "tests", "host_harness", "ba_nccl")
    if not os.path.exists(exe):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "host_harness")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_server_path_one_process_n_threads_over_nccl(tmp_path):
    """tests/host_harness/ba_nccl.cpp: N threads of ONE process, N ncclComm_t, the ncclAllReduce hook of INTEGRATION.md - what
    corbslam_server's GBA thread calls (GlobalOptimize.cpp:399,435-547). Its result equals the library's single-GPU solve."""
    import numpy as np
    import torch
    from corb_slam_b200 import Optimizer
    from corb_slam_b200.ba_file import read_result, write_problem
    from corb_slam_b200.synth import ba_problem
    prob = ba_problem(150, 12000, seed=7, n_fusion=10)
    P, L, E = write_problem(tmp_path / "p.bin", prob)
    full, finfo = Optimizer.BundleAdjustment(prob, 10, bRobust=False)
    for n in [1] + ([2] if torch.cuda.device_count() >= 2 else []) + ([4] if torch.cuda.device_count() >= 4 else []):
        r = subprocess.run([_harness(), str(n), str(tmp_path / "p.bin"), str(tmp_path / "r.bin"), "10"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        res = read_result(tmp_path / "r.bin", P, L)
        assert res["trial_accepted"] == finfo["trial_accepted"] and res["iterations"] == finfo["iterations"]
        assert res["rank_spread"] == 0.0
        assert np.abs(res["pose_t"] - full["pose_t"]).max() < 1e-7 and np.abs(res["point_xyz"] - full["point_xyz"]).max() < 1e-6
        if n == 1:
            assert np.array_equal(res["pose_t"], full["pose_t"])  # the same code path: bit-identical
