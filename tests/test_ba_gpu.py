"""GPU parity of the global bundle adjustment against the CPU oracle.

Tolerance (north_star): RMS reprojection error of the GPU and oracle solutions differs by < 1e-5 px, and the LM
accept/reject sequence is identical. Poses/points themselves are compared with a looser absolute tolerance because the
gauge is only fixed by one keyframe."""
import numpy as np
import pytest

import oracle
from oracle import _ba_bind as B
from corb_slam_b200 import Optimizer
from corb_slam_b200.synth import ba_problem

pytestmark = pytest.mark.gpu

RMS_TOL = 1e-5


def _check(prob, iters=10, robust=False, rms_tol=RMS_TOL):
    oracle.lib()
    o_out, o_info = B.solve(prob, iters, robust=robust)
    g_out, g_info = Optimizer.BundleAdjustment(prob, iters, bRobust=robust)
    assert g_info["trial_accepted"] == o_info["trial_accepted"]
    assert g_info["iterations"] == o_info["iterations"] and g_info["n_trials"] == o_info["n_trials"]
    assert g_info["lambda_initial"] == pytest.approx(o_info["lambda_initial"], rel=1e-9)
    assert g_info["chi2_initial"] == pytest.approx(o_info["chi2_initial"], rel=1e-10)
    assert g_info["chi2_final"] == pytest.approx(o_info["chi2_final"], rel=1e-6)
    _, o_rms = B.chi2(o_out)
    _, g_rms = B.chi2(g_out)
    assert abs(o_rms - g_rms) < rms_tol, (o_rms, g_rms)
    assert np.abs(g_out["point_xyz"] - o_out["point_xyz"]).max() < 1e-4
    assert np.abs(g_out["pose_t"] - o_out["pose_t"]).max() < 1e-4
    # the input arrays must not have been modified
    assert g_out["pose_q"] is not prob["pose_q"]
    return o_info, g_info, g_rms


@pytest.mark.parametrize("P,L", [(20, 2000), (200, 20000)])
def test_ba_matches_oracle(P, L):
    prob = ba_problem(P, L, seed=7, n_fusion=20)
    o_info, g_info, rms = _check(prob)
    assert o_info["chi2_final"] < 0.02 * o_info["chi2_initial"] and rms < 1.2


def test_ba_robust_huber():
    prob = ba_problem(30, 3000, seed=3, n_fusion=10)
    rng = np.random.default_rng(0)
    bad = rng.choice(len(prob["edge_pose"]), 150, replace=False)  # gross outliers
    prob["edge_obs"][bad, 0] += rng.normal(0, 40, len(bad))
    _check(prob, iters=8, robust=True)


def test_ba_fixed_vertices_and_mono_only():
    prob = ba_problem(25, 1500, seed=5, n_fusion=5, stereo_fraction=0.0)
    prob["pose_fixed"][:3] = 1                   # several anchors (client-side GBA fixes foreign keyframes, Cache.cc:482)
    prob["point_fixed"][::7] = 1                 # and foreign map points (Cache.cc:534)
    o_info, g_info, _ = _check(prob, iters=6)
    out, _ = Optimizer.BundleAdjustment(prob, 6, bRobust=False)
    np.testing.assert_array_equal(out["pose_t"][:3], prob["pose_t"][:3])
    np.testing.assert_array_equal(out["point_xyz"][::7], prob["point_xyz"][::7])


def test_ba_free_gauge_and_rejected_steps():
    # no fixed keyframe at all (server map without client 1): only the LM damping regularises
    prob = ba_problem(15, 800, seed=9, n_fusion=0)
    prob["pose_fixed"][:] = 0
    _check(prob, iters=5, rms_tol=1e-4)
    # a badly perturbed start (every trial is still accepted: Gauss-Newton steps on a consistent problem keep reducing chi2)
    prob = ba_problem(15, 800, seed=11, n_fusion=0)
    rng = np.random.default_rng(1)
    prob["point_xyz"] += rng.normal(0, 1.5, prob["point_xyz"].shape)
    prob["pose_t"][1:] += rng.normal(0, 0.8, prob["pose_t"][1:].shape)
    o_info, g_info, _ = _check(prob, iters=10, rms_tol=1e-3)


def _with_gross_outliers(prob, seed, fraction=0.2, sigma=300.0):
    """Wrong data associations: `fraction` of the observations are off by hundreds of pixels. The problem is inconsistent, chi2
    is noise dominated and LM REJECTS trials (rho < 0, lambda *= nu, nu *= 2) - the only way found to make the
    accept/reject gate non-trivial (perturbing the start alone never produced a rejection, at any size)."""
    p = dict(prob)
    p["edge_obs"] = prob["edge_obs"].copy()
    rng = np.random.default_rng(seed)
    bad = rng.choice(len(p["edge_obs"]), int(len(p["edge_obs"]) * fraction), replace=False)
    p["edge_obs"][bad, :2] += rng.normal(0, sigma, (len(bad), 2))
    return p


@pytest.mark.parametrize("stereo_fraction,robust", [(0.0, False), (0.8, False), (0.8, True)])
def test_ba_rejected_trials_sequence_is_identical(stereo_fraction, robust):
    prob = _with_gross_outliers(ba_problem(60, 3000, seed=11, n_fusion=0, stereo_fraction=stereo_fraction), 1)
    o_info, g_info, _ = _check(prob, iters=10, robust=robust, rms_tol=1e-3)
    assert 0 in o_info["trial_accepted"] and 1 in o_info["trial_accepted"] and o_info["n_trials"] > o_info["iterations"]


def test_ba_headline_size_with_rejected_trials():
    """The bench problem (P = 2000 keyframes, L = 200 k landmarks, ~1 M observations, BASELINE.json config 5) against the
    oracle: once as the bench runs it, once with gross outliers so that trials are rejected at this size too."""
    prob = ba_problem(2000, 200000, seed=7)
    o_info, g_info, rms = _check(prob, iters=10)
    assert rms < 1.1 and g_info["band_chunks"] > 1
    o_info, g_info, _ = _check(_with_gross_outliers(prob, 3), iters=10, rms_tol=1e-3)
    assert 0 in o_info["trial_accepted"] and o_info["n_trials"] > o_info["iterations"]


def test_ba_stop_flag_and_degenerate_inputs():
    prob = ba_problem(20, 1000, seed=2)
    stop = np.ones(1, np.uint8)
    out, info = Optimizer.BundleAdjustment(prob, 10, pbStopFlag=stop, bRobust=False)
    assert info["iterations"] == 0 and info["stopped"] == 1
    np.testing.assert_array_equal(out["point_xyz"], prob["point_xyz"])
    # zero iterations / points without edges / an empty problem
    out, info = Optimizer.BundleAdjustment(prob, 0, bRobust=False)
    assert info["iterations"] == 0
    lonely = dict(prob)
    lonely["point_xyz"] = np.vstack([prob["point_xyz"], [[1.0, 2.0, 30.0]]])
    lonely["point_fixed"] = np.append(prob["point_fixed"], 0).astype(np.uint8)
    o_out, o_info = B.solve(lonely, 4)
    g_out, g_info = Optimizer.BundleAdjustment(lonely, 4, bRobust=False)
    assert g_info["trial_accepted"] == o_info["trial_accepted"]
    np.testing.assert_array_equal(g_out["point_xyz"][-1], [1.0, 2.0, 30.0])
    with pytest.raises(Exception):
        bad = dict(prob)
        bad["edge_pose"] = prob["edge_pose"].copy()
        bad["edge_pose"][0] = 10 ** 6
        Optimizer.BundleAdjustment(bad, 1, bRobust=False)


def test_ba_edge_order_does_not_matter():
    """Edges grouped by landmark are used in place (the order Optimizer.cc:131-196 creates them in); any other order is
    stably sorted by landmark first. Both routes must give the same optimisation."""
    prob = ba_problem(40, 4000, seed=13, n_fusion=8)
    out_a, info_a = Optimizer.BundleAdjustment(prob, 6, bRobust=False)
    rng = np.random.default_rng(4)
    perm = rng.permutation(len(prob["edge_pose"]))
    shuf = dict(prob)
    for k in ("edge_pose", "edge_point", "edge_obs", "edge_inv_sigma2"):
        shuf[k] = np.ascontiguousarray(prob[k][perm])
    out_b, info_b = Optimizer.BundleAdjustment(shuf, 6, bRobust=False)
    assert info_a["trial_accepted"] == info_b["trial_accepted"]
    assert info_b["chi2_final"] == pytest.approx(info_a["chi2_final"], rel=1e-9)
    np.testing.assert_allclose(out_b["pose_t"], out_a["pose_t"], atol=1e-8)
    np.testing.assert_allclose(out_b["point_xyz"], out_a["point_xyz"], atol=1e-8)


def test_ba_chunked_band_equals_single_sweep(monkeypatch):
    """Long trajectories are cut into independent band chunks (separator keyframes join the border block); the result
    must agree with the single sweep (CORB_BA_CHUNKS=1) and with the oracle."""
    prob = ba_problem(600, 20000, seed=17, n_fusion=12)
    monkeypatch.setenv("CORB_BA_CHUNKS", "1")
    out_1, info_1 = Optimizer.BundleAdjustment(prob, 5, bRobust=False)
    monkeypatch.delenv("CORB_BA_CHUNKS")
    out_q, info_q = Optimizer.BundleAdjustment(prob, 5, bRobust=False)
    assert info_1["band_chunks"] == 1 and info_q["band_chunks"] > 1 and info_q["separator_poses"] > 0
    assert info_q["trial_accepted"] == info_1["trial_accepted"]
    assert info_q["chi2_final"] == pytest.approx(info_1["chi2_final"], rel=1e-9)
    np.testing.assert_allclose(out_q["pose_t"], out_1["pose_t"], atol=1e-7)
    _check(prob, iters=5)


def test_ba_dense_covisibility():
    """Landmarks seen by 40 consecutive keyframes: a reduced-system row then has more distinct blocks than the per-warp
    and per-row caches of k_ba_schur_rows hold (single-warp fallback with global accumulation), and the band is too wide
    to be cut into chunks."""
    prob = ba_problem(70, 2500, seed=21, obs_per_point=40, n_fusion=6)
    o_info, g_info, _ = _check(prob, iters=5)
    assert g_info["band_chunks"] == 1
    prob = ba_problem(60, 3000, seed=22, obs_per_point=20, n_fusion=0)   # fits the merged cache (<= 32 blocks), not one warp's
    _check(prob, iters=5)


def test_ba_arena_release_and_reuse():
    """corb_ba_release_cache returns the per-device arena; the next call rebuilds it and gives the same result."""
    from corb_slam_b200.optimizer import release_cache
    prob = ba_problem(30, 3000, seed=4, n_fusion=4)
    a, ia = Optimizer.BundleAdjustment(prob, 4, bRobust=False)
    release_cache(0)
    release_cache(0)
    b, ib = Optimizer.BundleAdjustment(prob, 4, bRobust=False)
    assert ia["chi2_final"] == ib["chi2_final"]
    np.testing.assert_array_equal(a["pose_t"], b["pose_t"])


def test_ba_large_problem_staged_upload_equals_plain_upload(monkeypatch):
    """Above ~420 k observations the edge arrays travel through page-locked staging slots on helper threads (part of the
    per-device arena); without the arena they are plain copies. Same bits either way, and the chi2 the oracle reports for
    the GPU solution equals the library's own."""
    prob = ba_problem(500, 100000, seed=11, n_fusion=6)
    assert len(prob["edge_pose"]) * 40 >= 16 << 20
    a, ia = Optimizer.BundleAdjustment(prob, 3, bRobust=False)
    monkeypatch.setenv("CORB_BA_NO_ARENA", "1")
    b, ib = Optimizer.BundleAdjustment(prob, 3, bRobust=False)
    monkeypatch.delenv("CORB_BA_NO_ARENA")
    assert ia["chi2_final"] == ib["chi2_final"] and ia["trial_accepted"] == ib["trial_accepted"]
    np.testing.assert_array_equal(a["pose_t"], b["pose_t"])
    np.testing.assert_array_equal(a["point_xyz"], b["point_xyz"])
    # shuffled edges (not grouped by landmark): the library sorts them, the staged path uploads the sorted copies
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(prob["edge_pose"]))
    shuf = dict(prob)
    for k in ("edge_pose", "edge_point", "edge_obs", "edge_inv_sigma2"):
        shuf[k] = np.ascontiguousarray(prob[k][perm])
    c, ic = Optimizer.BundleAdjustment(shuf, 3, bRobust=False)
    assert ic["trial_accepted"] == ia["trial_accepted"] and ic["chi2_final"] == pytest.approx(ia["chi2_final"], rel=1e-9)
    oracle.lib()
    chi2, _ = B.chi2(a)
    assert chi2 == pytest.approx(ia["chi2_final"], rel=1e-9)
    with pytest.raises(Exception):  # an out-of-range edge anywhere in the list is reported, nothing is computed
        bad = dict(prob)
        bad["edge_pose"] = prob["edge_pose"].copy()
        bad["edge_pose"][len(perm) // 2] = 500
        Optimizer.BundleAdjustment(bad, 1, bRobust=False)


def test_ba_schur_pair_lists_equal_the_edge_walk(monkeypatch):
    """The Schur rows from the sorted pair lists (k_ba_schur_pairs, built once per call) against the per-iteration edge walk
    (k_ba_schur_rows, CORB_BA_SCHUR=old): same accept / reject sequence, chi2 to 1e-10, poses to 1e-8 - with fixed keyframes,
    fixed landmarks, fusion links and both kernel shapes; and the pair lists are what runs by default, reproducibly bit for bit."""
    prob = ba_problem(600, 20000, seed=23, n_fusion=12)
    prob["pose_fixed"][:2] = 1
    prob["point_fixed"][::11] = 1
    monkeypatch.setenv("CORB_BA_SCHUR", "old")
    out_o, info_o = Optimizer.BundleAdjustment(prob, 5, bRobust=False)
    monkeypatch.delenv("CORB_BA_SCHUR")
    assert info_o["schur_pair_lists"] == 0
    out_n, info_n = Optimizer.BundleAdjustment(prob, 5, bRobust=False)
    out_n2, info_n2 = Optimizer.BundleAdjustment(prob, 5, bRobust=False)
    assert info_n["schur_pair_lists"] == 1
    assert info_n["trial_accepted"] == info_o["trial_accepted"]
    assert info_n["chi2_final"] == pytest.approx(info_o["chi2_final"], rel=1e-10)
    np.testing.assert_allclose(out_n["pose_t"], out_o["pose_t"], atol=1e-8)
    np.testing.assert_allclose(out_n["point_xyz"], out_o["point_xyz"], atol=1e-7)
    assert out_n["pose_t"].tobytes() == out_n2["pose_t"].tobytes() and info_n["chi2_final"] == info_n2["chi2_final"]  # fixed summation order
    monkeypatch.setenv("CORB_BA_SP", "256,64")
    out_s, info_s = Optimizer.BundleAdjustment(prob, 5, bRobust=False)
    monkeypatch.delenv("CORB_BA_SP")
    assert info_s["schur_pair_lists"] == 1 and info_s["trial_accepted"] == info_o["trial_accepted"]
    assert info_s["chi2_final"] == pytest.approx(info_o["chi2_final"], rel=1e-10)
    _check(prob, iters=5)


def test_ba_page_locked_edge_arrays_equal_pageable():
    """Edge arrays in page-locked memory (corb_host_alloc: what the C++ shim flattens into) are uploaded by the DMA engines in
    place, pageable ones through the staging slots: same bits either way."""
    from corb_slam_b200 import page_locked
    prob = ba_problem(400, 30000, seed=29, n_fusion=6)
    out_a, info_a = Optimizer.BundleAdjustment(prob, 4, bRobust=False)
    out_b, info_b = Optimizer.BundleAdjustment(page_locked(prob), 4, bRobust=False)
    assert info_a["trial_accepted"] == info_b["trial_accepted"] and info_a["chi2_final"] == info_b["chi2_final"]
    assert out_a["pose_t"].tobytes() == out_b["pose_t"].tobytes() and out_a["point_xyz"].tobytes() == out_b["point_xyz"].tobytes()
