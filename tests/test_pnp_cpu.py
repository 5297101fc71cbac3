"""EPnP-RANSAC, CPU side: the oracle (oracle/pnp_oracle.cpp) against the golden vectors of the real
OpenCV (tests/golden/pnp_cv2.npz, tools/gen_golden_pnp.py), its RANSAC bookkeeping, the library's SetRansacParameters, and the
per-thread EPnP of the CUDA path compiled for the host (tests/host_harness) bit for bit against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle import _pnp_bind as P
from corb_slam_b200.synth import pnp_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "pnp_cv2.npz"))


@pytest.fixture(scope="module", autouse=True)
def _oracle():
    oracle.lib()


def _vec_close(a, b, tol):
    return min(np.abs(a - b).max(), np.abs(a + b).max()) < tol  # singular vectors are defined up to sign


def test_svd_solve_invert_match_opencv():
    for si, (m, n) in enumerate([(12, 12), (3, 3), (6, 4), (6, 3), (6, 5)]):
        for rep in range(4):
            k = "svd_%d_%d" % (si, rep)
            A = G[k + "_A"]
            Ut, W, Vt = P.svd(A)
            assert np.abs(W - G[k + "_w"]).max() <= 1e-12 * G[k + "_w"].max()
            assert np.all(np.diff(W) <= 0)
            for i in range(n):
                assert _vec_close(Ut[i], G[k + "_u"][:, i], 1e-9) and _vec_close(Vt[i], G[k + "_vt"][i], 1e-9), (k, i)
            assert np.abs((Ut.T * W) @ Vt - A).max() < 1e-12 * max(1.0, np.abs(A).max())
            if m <= 6:
                assert np.abs(P.svd_solve(A, G[k + "_b"]) - G[k + "_x"]).max() < 1e-11
            if m == 3:
                assert np.abs(P.svd_invert3(A) - G[k + "_inv"]).max() < 1e-10 * np.abs(G[k + "_inv"]).max()


def test_epnp_matches_opencv_epnp():
    """compute_pose against cv2.solvePnP(SOLVEPNP_EPNP) - OpenCV's copy of the EPnP code the reference vendors."""
    for ci in range(int(G["n_epnp"])):
        k = "epnp_%d" % ci
        fx, fy, cx, cy = G[k + "_K"]
        R, t, err = P.epnp_pose(G[k + "_Xw"], G[k + "_uv"], fx, fy, cx, cy)
        assert np.abs(R - G[k + "_R"]).max() < 1e-9 and np.abs(t - G[k + "_t"]).max() < 1e-8, (ci, np.abs(R - G[k + "_R"]).max())
        assert abs(np.linalg.det(R) - 1) < 1e-9
        if ci % 2 == 0:  # noise-free: the true pose (float32 inputs)
            assert np.abs(R - G[k + "_Rtrue"]).max() < 1e-4 and np.abs(t - G[k + "_ttrue"]).max() < 2e-3


def _draws(rng, n, its):
    return np.stack([rng.integers(0, n - k, its) for k in range(4)], 1).astype(np.int32)


def _oracle_solver(p, min_inl=10, max_it=300, eps=0.5, th2=5.991):
    n = len(p["p2d"])
    mi, its = P.ransac_params(n, 0.99, min_inl, max_it, 4, eps)
    max_err = (p["sigma2"] * np.float32(th2)).astype(np.float32)
    return P.PnpSolver(p["p2d"], p["p3d"], max_err, *[float(v) for v in p["K"]], mi, its), mi, its


def test_oracle_ransac_recovers_the_pose_and_keeps_state():
    rng = np.random.default_rng(0)
    found = 0
    for seed in range(12):
        p = pnp_problem(seed, n=150, outlier_fraction=0.25, pixel_noise=0.5)
        s, mi, its = _oracle_solver(p)
        d = _draws(rng, 150, its + 20)
        rc, no_more, inl, n_inl, T = s.iterate(5, d)
        assert rc in (0, 1, 2)
        if rc == 1:
            found += 1
            assert not no_more and n_inl == inl.sum() > mi and s.iterations <= its
            assert np.abs(T[:3, :3] - p["Tcw"][:, :3]).max() < 5e-3 and np.abs(T[:3, 3] - p["Tcw"][:, 3]).max() < 0.1
            assert (inl & p["outlier"]).sum() <= 3
            # the next call resumes where this one returned (state kept in the object)
            done = s.iterations
            rc2, _, inl2, n2, T2 = s.iterate(5, d)
            assert s.iterations > done and (rc2 != 1 or n2 > mi)
        else:
            assert no_more and s.iterations == its
    assert found >= 11


def test_oracle_ransac_edge_cases():
    rng = np.random.default_rng(1)
    p = pnp_problem(3, n=8, outlier_fraction=0.0)
    s, mi, its = _oracle_solver(p)                       # N < mRansacMinInliers: bNoMore without iterating (:215-219)
    assert mi == 10
    rc, no_more, inl, n_inl, T = s.iterate(5, _draws(rng, 8, 10))
    assert rc == 0 and no_more and s.iterations == 0 and n_inl == 0
    p = pnp_problem(4, n=60, outlier_fraction=1.0)       # no consistent pose: exhausted, empty or best-at-end
    s, mi, its = _oracle_solver(p)
    rc, no_more, inl, n_inl, T = s.iterate(5, _draws(rng, 60, its))
    assert rc in (0, 2) and no_more and s.iterations == its
    with pytest.raises(ValueError):
        _oracle_solver(pnp_problem(5, n=60))[0].iterate(5, _draws(rng, 60, 2))  # too few draws


def test_ransac_params_library_equals_oracle():
    from corb_slam_b200 import _lib
    L = _lib.lib()
    for N in list(range(0, 40)) + [63, 64, 100, 150, 500, 2000]:
        for (prob, mi, it, ms, eps) in [(0.99, 10, 300, 4, 0.5), (0.99, 20, 300, 4, 0.4), (0.99, 8, 300, 4, 0.4), (0.999, 15, 50, 4, 0.3)]:
            a, b = C.c_int(0), C.c_int(0)
            assert L.corb_pnp_ransac_params(N, prob, mi, it, ms, eps, C.byref(a), C.byref(b)) == 0
            if N > 0:
                assert (a.value, b.value) == P.ransac_params(N, prob, mi, it, ms, eps), N
            assert a.value >= 4 and 1 <= b.value <= it
    # the reference's own settings (Tracking.cc:1414): 150 matches -> min inliers 75, ceil(log(0.01)/log(1-0.125)) = 35 iterations
    assert P.ransac_params(150, 0.99, 10, 300, 4, 0.5) == (75, 35)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pnp") / "pnp_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", so,
                    os.path.join(ROOT, "tests", "host_harness", "pnp_host.cpp")], check=True)
    H = C.CDLL(so)
    vp = C.c_void_p
    H.harness_epnp_pose.restype = C.c_double
    H.harness_epnp_pose.argtypes = [vp, vp, C.c_int, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, vp]
    H.harness_count_inliers.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp, vp, C.c_int, vp]
    return H


def test_device_epnp_compiled_for_the_host_is_bit_exact(harness):
    """pnp_core.cuh (what every GPU thread runs, built there with -fmad=false) against the oracle: same bits, for minimal
    sets - where M^T M has a 4-dimensional null space and only an identical operation sequence can agree - and for masks."""
    rng = np.random.default_rng(0)
    for seed in range(60):
        n = 40 + 7 * seed
        p = pnp_problem(seed, n=n)
        fx, fy, cx, cy = [float(v) for v in p["K"]]
        lst = rng.choice(n, 4, replace=False).astype(np.int32)
        Rt = np.zeros(12)
        e = harness.harness_epnp_pose(p["p3d"].ctypes.data, p["p2d"].ctypes.data, n, lst.ctypes.data, None, fx, fy, cx, cy,
                                      1 + 31 * (seed & 1), Rt.ctypes.data)
        R, t, eo = P.epnp_pose(p["p3d"][lst].astype(np.float64), p["p2d"][lst].astype(np.float64), fx, fy, cx, cy)
        assert Rt[:9].tobytes() == R.tobytes() and Rt[9:].tobytes() == t.tobytes() and e == eo, seed
        sel = ~p["outlier"]
        mask = np.packbits(np.r_[sel, np.zeros((-n) % 32, bool)], bitorder="little").view(np.uint32)
        e = harness.harness_epnp_pose(p["p3d"].ctypes.data, p["p2d"].ctypes.data, n, None, mask.ctypes.data, fx, fy, cx, cy, 32,
                                      Rt.ctypes.data)
        R, t, eo = P.epnp_pose(p["p3d"][sel].astype(np.float64), p["p2d"][sel].astype(np.float64), fx, fy, cx, cy)
        assert Rt[:9].tobytes() == R.tobytes() and Rt[9:].tobytes() == t.tobytes() and e == eo, seed


def test_draw_resolution_equals_the_list_bookkeeping(harness):
    """resolve_draws (the kernel's replay of vAvailableIndices, PnPsolver.cc:228-242) against the literal list operations,
    exhaustively for small n and on random draws for large n."""
    def model(n, r):
        avail = list(range(n))
        out = []
        for k in range(4):
            out.append(avail[r[k]])
            avail[r[k]] = avail[-1]
            avail.pop()
        return out

    def dev(n, r):
        rr = np.array(r, np.int32)
        out = np.zeros(4, np.int32)
        harness.harness_resolve_draws(n, rr.ctypes.data, out.ctypes.data)
        return out.tolist()

    harness.harness_resolve_draws.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    harness.harness_resolve_draws.restype = None
    for n in (4, 5, 6, 7):
        for a in range(n):
            for b in range(n - 1):
                for c in range(n - 2):
                    for d in range(n - 3):
                        assert dev(n, (a, b, c, d)) == model(n, (a, b, c, d)), (n, a, b, c, d)
    rng = np.random.default_rng(0)
    for _ in range(2000):
        n = int(rng.integers(4, 3000))
        r = [int(rng.integers(0, n - k)) for k in range(4)]
        got = dev(n, r)
        assert got == model(n, r) and len(set(got)) == 4


def test_wavefront_schedule_preserves_the_order_of_dependent_visits():
    """jacobi12_wavefront runs visit (i, j) of a cyclic-by-rows sweep at step i + j. The claim behind its bit-exactness: visits
    of one step touch disjoint rows, and two visits that share a row keep their lexicographic (= sequential) order."""
    n = 12
    visits = [(i, j) for i in range(n - 1) for j in range(i + 1, n)]
    step = {v: v[0] + v[1] for v in visits}
    by_step = {}
    for v in visits:
        by_step.setdefault(step[v], []).append(v)
    assert sorted(by_step) == list(range(1, 2 * n - 2)) and max(len(b) for b in by_step.values()) == n // 2
    for vs in by_step.values():
        rows = [r for v in vs for r in v]
        assert len(rows) == len(set(rows))
        # the kernel's lane -> visit map: i = max(0, t - 11) + lane, j = t - i, active while i < j
        t = vs[0][0] + vs[0][1]
        lanes = [(max(0, t - (n - 1)) + l, t - (max(0, t - (n - 1)) + l)) for l in range(8)]
        assert sorted(v for v in lanes if v[0] < v[1]) == sorted(vs)
    for a in range(len(visits)):
        for b in range(a + 1, len(visits)):
            if set(visits[a]) & set(visits[b]):
                assert step[visits[a]] < step[visits[b]], (visits[a], visits[b])
