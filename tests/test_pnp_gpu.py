"""EPnP-RANSAC on the GPU (corb_pnp_iterate_batch through the PnPsolver mirror) against the sequential oracle on the same
draws: status, bNoMore, iteration count, inlier sets and the float32 Tcw bit for bit (pnp.cu is compiled with -fmad=false, so a
GPU thread performs the oracle's operation sequence)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import _pnp_bind as P
from corb_slam_b200 import CorbError, PnPsolver, _lib
from corb_slam_b200.synth import pnp_problem

pytestmark = pytest.mark.gpu


def _draws(rng, n, its):
    return np.stack([rng.integers(0, n - k, its) for k in range(4)], 1).astype(np.int32)


def _pair(p, draws, params=(0.99, 10, 300, 4, 0.5, 5.991)):
    n = len(p["p2d"])
    g = PnPsolver(p["p2d"], np.arange(n) % 8, np.float32(1.2) ** (2 * np.arange(8, dtype=np.float32)), p["K"], p["p3d"], np.ones(n, bool))
    g.mvSigma2 = p["sigma2"]
    g.SetRansacParameters(*params)
    g.set_draws(draws)
    o = P.PnpSolver(p["p2d"], p["p3d"], g.mvMaxError, *[float(v) for v in p["K"]], g.mRansacMinInliers, g.mRansacMaxIts)
    return g, o


def _same(res, ores, g, o, tag):
    Tcw, no_more, inl, n_inl = res
    rc, ono_more, oinl, on_inl, oT = ores
    assert (Tcw is not None) == (rc != 0), tag
    assert no_more == ono_more and n_inl == on_inl and g.mnIterations == o.iterations, (tag, no_more, ono_more, n_inl, on_inl, g.mnIterations, o.iterations)
    assert np.array_equal(inl, oinl), tag
    if rc:
        assert Tcw.tobytes() == oT.tobytes(), (tag, np.abs(Tcw - oT).max())


@pytest.mark.parametrize("team", ["1", "0"])
def test_batch_equals_sequential_oracle(monkeypatch, team):
    """team=1: a team of 8 lanes per hypothesis (wavefront Jacobi); team=0: one thread per hypothesis. Same bits either way."""
    monkeypatch.setenv("CORB_PNP_TEAM", team)
    oracle.lib()
    rng = np.random.default_rng(0)
    gs, os_, ds = [], [], []
    for c in range(24):
        n = [150, 60, 333, 31, 1000, 97][c % 6]
        p = pnp_problem(100 + c, n=n, outlier_fraction=[0.25, 0.45, 0.1, 0.97][c % 4], pixel_noise=0.5)
        d = _draws(rng, n, 400)
        g, o = _pair(p, d)
        gs.append(g); os_.append(o); ds.append(d)
    res = PnPsolver.iterate_batch(gs, 5)
    kinds = set()
    for c, (g, o) in enumerate(zip(gs, os_)):
        ores = o.iterate(5, ds[c])
        kinds.add(ores[0])
        _same(res[c], ores, g, o, c)
    assert kinds == {0, 1, 2} or kinds == {0, 1}, kinds
    # second round: every solver resumes from its own mnIterations (those that returned a pose continue, :1438 loop)
    res = PnPsolver.iterate_batch(gs, 5)
    for c, (g, o) in enumerate(zip(gs, os_)):
        _same(res[c], o.iterate(5, ds[c]), g, o, ("resume", c))
    # and single-solver calls give the same as the batch
    g, o = _pair(pnp_problem(100, n=150, outlier_fraction=0.25, pixel_noise=0.5), ds[0])
    _same(g.iterate(5), o.iterate(5, ds[0]), g, o, "single")


def test_pose_accuracy_and_find():
    rng = np.random.default_rng(2)
    ok = 0
    for seed in range(10):
        p = pnp_problem(seed, n=200, outlier_fraction=0.2, pixel_noise=0.3)
        g, o = _pair(p, _draws(rng, 200, 400))
        Tcw, inl, n_inl = g.find()
        if Tcw is not None:
            ok += 1
            assert np.abs(Tcw[:3, :3] - p["Tcw"][:, :3]).max() < 3e-3 and np.abs(Tcw[:3, 3] - p["Tcw"][:, 3]).max() < 0.1
            assert n_inl == inl.sum() and (inl & p["outlier"]).sum() <= 3
    assert ok >= 9


def test_draw_resolution_corner_cases():
    """vAvailableIndices bookkeeping (:228-242): repeated positions, the last position, descending picks."""
    oracle.lib()
    p = pnp_problem(7, n=40, outlier_fraction=0.0, pixel_noise=0.2)
    n = 40
    rows = [[0, 0, 0, 0], [39, 38, 37, 36], [39, 0, 37, 0], [38, 38, 37, 36], [5, 38, 5, 36], [0, 38, 0, 0], [36, 36, 36, 36], [39, 38, 0, 0],
            [1, 1, 37, 1], [38, 0, 37, 36]]
    rng = np.random.default_rng(3)
    d = np.array(rows + _draws(rng, n, 60).tolist(), np.int32)
    for start in range(len(rows)):
        g, o = _pair(p, d[start:], params=(0.99, 38, 300, 4, 0.5, 5.991))  # min inliers 38 of 40: most hypotheses fail -> many are visited
        _same(g.iterate(5), o.iterate(5, d[start:]), g, o, start)


def test_edge_cases():
    oracle.lib()
    rng = np.random.default_rng(4)
    # N < mRansacMinInliers: bNoMore, nothing iterated; mixed into a batch with a live problem
    small, _ = _pair(pnp_problem(1, n=8, outlier_fraction=0.0), _draws(rng, 8, 10))
    d = _draws(rng, 64, 400)
    live, o = _pair(pnp_problem(2, n=64, outlier_fraction=0.2, pixel_noise=0.4), d)
    r = PnPsolver.iterate_batch([small, live], 5)
    assert r[0][0] is None and r[0][1] and small.mnIterations == 0 and r[0][3] == 0
    _same(r[1], o.iterate(5, d), live, o, "live")
    # all outliers: the iterations are exhausted
    d = _draws(rng, 80, 400)
    g, o = _pair(pnp_problem(3, n=80, outlier_fraction=1.0), d)
    res = g.iterate(5)
    _same(res, o.iterate(5, d), g, o, "outliers")
    assert res[1] and g.mnIterations == g.mRansacMaxIts
    # exhausted solver called again: nIterations more iterations (the || of the loop condition, :223)
    _same(g.iterate(5), o.iterate(5, d), g, o, "again")
    assert g.mnIterations == g.mRansacMaxIts + 5
    # mRansacMinInliers == N -> one iteration (:184-185); masked matches (mvKeyPointIndices scatter)
    p = pnp_problem(5, n=12, outlier_fraction=0.0, pixel_noise=0.1)
    valid = np.ones(20, bool); valid[[3, 9, 10, 15, 16, 17, 18, 19]] = False
    keys = np.zeros((20, 2), np.float32); keys[valid] = p["p2d"]
    xyz = np.zeros((20, 3), np.float32); xyz[valid] = p["p3d"]
    g = PnPsolver(keys, np.zeros(20, int), [1.0], p["K"], xyz, valid)
    g.SetRansacParameters(0.99, 12, 300, 4, 0.5, 5.991)
    assert g.N == 12 and g.mRansacMinInliers == 12 and g.mRansacMaxIts == 1
    g.set_draws(_draws(rng, 12, 10))
    Tcw, no_more, inl, n_inl = g.iterate(5)
    assert len(inl) == 20 and not inl[~valid].any() and g.mnIterations >= 1
    # a draw outside [0, N-k) is rejected, nothing is computed
    g, o = _pair(pnp_problem(6, n=50), _draws(rng, 50, 400))
    bad = g._draws.copy(); bad[3, 2] = 48
    g.set_draws(bad)
    with pytest.raises(CorbError) as e:
        g.iterate(5)
    assert e.value.status == _lib.ERR_INVALID


def test_many_candidates_large():
    """Map-fusion sized batch: 40 candidates x 2000 matches; agreement with the oracle on every candidate."""
    oracle.lib()
    rng = np.random.default_rng(5)
    gs, os_, ds = [], [], []
    for c in range(40):
        n = 2000 if c % 5 == 0 else 300
        p = pnp_problem(300 + c, n=n, outlier_fraction=0.3 if c % 3 else 0.9, pixel_noise=0.5)
        d = _draws(rng, n, 320)
        g, o = _pair(p, d, params=(0.99, 20, 300, 4, 0.4, 5.991))  # LoopClosing.cc:277-style settings
        gs.append(g); os_.append(o); ds.append(d)
    res = PnPsolver.iterate_batch(gs, 5)
    for c, (g, o) in enumerate(zip(gs, os_)):
        _same(res[c], o.iterate(5, ds[c]), g, o, c)
