"""CPU suite: pins the oracle's OpenCV-primitive models against golden vectors produced by the real cv2 4.13.0
(tests/golden/opencv_primitives.npz, made by tools/gen_golden.py), checks the reference's source constants and the
composed extractor's invariants, and the matching / BoW / BA restatements by algebra."""
import os

import numpy as np
import pytest

import oracle
from oracle import _ba_bind as B
from oracle import _match_bind as M
from corb_slam_b200.synth import ba_problem, ba_shard, stereo_frame

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    oracle.lib()
    return np.load(os.path.join(GOLD, "opencv_primitives.npz"))


# ------------------------------------------------------------------------------------------------ OpenCV primitives
def test_resize_linear_matches_cv2_golden(gold):
    for name, sizes in (("scene", [(267, 167), (222, 139), (185, 116)]), ("noise", [(109, 81), (91, 67), (40, 23)])):
        cur = gold["img_" + name]
        for i, (w, h) in enumerate(sizes):
            got = oracle.resize_linear(cur, w, h)
            np.testing.assert_array_equal(got, gold["resize_%s_%d" % (name, i)])
            cur = got


def test_gaussian7_matches_cv2_golden(gold):
    for name in ("scene", "noise", "tiny"):
        np.testing.assert_array_equal(oracle.gaussian7(gold["img_" + name]), gold["blur_" + name])


def test_fast_matches_cv2_golden(gold):
    for name in ("scene", "noise", "cell"):
        for th in (20, 7):
            got = oracle.fast_detect(gold["img_" + name], th)
            np.testing.assert_array_equal(got, gold["fast_%s_%d" % (name, th)])  # order, position and response
            # the score map alone reproduces the detector: corner at th <=> score >= th, strict 8-neighbour maximum
            sc = oracle.fast_score(gold["img_" + name]).astype(np.int32)
            m = np.where(sc >= th, sc, 0)
            keep = []
            for y, x in zip(*np.nonzero(m)):
                nb = m[y - 1:y + 2, x - 1:x + 2].copy()
                nb[1, 1] = -1
                if (m[y, x] > nb).all():
                    keep.append((x, y, m[y, x]))
            np.testing.assert_array_equal(np.array(keep, np.int32).reshape(-1, 3), got)


def test_fast_atan2_matches_cv2_golden(gold):
    got = np.array([oracle.fast_atan2(y, x) for y, x in zip(gold["atan2_y"], gold["atan2_x"])], np.float32)
    np.testing.assert_array_equal(got.view(np.uint32), gold["atan2_deg"].view(np.uint32))


def test_cv_round_half_to_even():
    assert [oracle.cv_round(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5, 2.4999, 2.5001)] == [0, 2, 2, 0, -2, 2, 3]


def test_live_cv2_if_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (75, 113), dtype=np.uint8)
    np.testing.assert_array_equal(oracle.resize_linear(img, 94, 62), cv2.resize(img, (94, 62), interpolation=cv2.INTER_LINEAR))
    np.testing.assert_array_equal(oracle.gaussian7(img), cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101))


# ------------------------------------------------------------------------------------------------ extractor
def test_constructor_tables_match_survey_appendix_a():
    ex = oracle.OrbExtractor(2000, 1.2, 8, 20, 7)
    assert list(ex.quota) == [434, 362, 302, 251, 209, 175, 145, 122] and ex.quota.sum() == 2000
    assert list(ex.umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]  # ORBextractor.cc:452-469
    assert np.float32(ex.scale[6]) == np.float32(2.9859846) and np.float32(ex.scale[7]) == np.float32(3.5831816)
    table = {(1242, 375): [(1242, 375), (1035, 312), (862, 260), (719, 217), (599, 181), (499, 151), (416, 126), (347, 105)],
             (1241, 376): [(1241, 376), (1034, 313), (862, 261), (718, 218), (598, 181), (499, 151), (416, 126), (346, 105)],
             (1226, 370): [(1226, 370), (1022, 308), (851, 257), (709, 214), (591, 178), (493, 149), (411, 124), (342, 103)]}
    for (w, h), levels in table.items():
        assert [ex.level_size(l, w, h) for l in range(8)] == levels
    assert sum(a * b for a, b in table[(1242, 375)]) == 1441432
    assert ex.capacity(1242, 375) == 2000 + 3 * 8
    assert ex.capacity(40, 40) < 0  # smaller than the 30 px FAST grid: the reference cannot process it either


def test_extractor_invariants_and_regression():
    left, right = stereo_frame(1234, w=640, h=240)
    ex = oracle.OrbExtractor(500, 1.2, 4, 20, 7)
    kps, desc = ex(left)
    ref = np.load(os.path.join(GOLD, "oracle_extract_640x240.npz"))
    assert kps.tobytes() == ref["kps"].tobytes()
    np.testing.assert_array_equal(desc, ref["desc"])
    kps2, desc2 = ex(left)
    assert kps2.tobytes() == kps.tobytes()  # deterministic, no state carried between frames
    for l in range(4):
        n, q = ex.level_count(l), ex.quota[l]
        assert q <= n <= q + 2 or n == len(ex.candidates(l))  # DistributeOctTree overshoots by at most 2
        lw, lh = ex.level_size(l, 640, 240)
        k = kps[kps["octave"] == l]
        x = k["x"] / ex.scale[l] if l else k["x"]
        y = k["y"] / ex.scale[l] if l else k["y"]
        assert (np.rint(x) >= 19).all() and (np.rint(x) < lw - 19).all() and (np.rint(y) >= 19).all() and (np.rint(y) < lh - 19).all()
        assert (k["size"] == np.float32(int(31 * ex.scale[l]))).all()
    assert (kps["class_id"] == -1).all() and (kps["angle"] >= 0).all() and (kps["angle"] < 360).all()
    # stage taps agree with the primitives
    np.testing.assert_array_equal(ex.pyramid(1), oracle.resize_linear(left, *ex.level_size(1, 640, 240)))
    np.testing.assert_array_equal(ex.blurred(1), oracle.gaussian7(ex.pyramid(1)))
    k0 = kps[0]
    assert oracle.ic_angle(ex.pyramid(0), int(k0["x"]), int(k0["y"]), ex.umax) == k0["angle"]
    np.testing.assert_array_equal(oracle.brief(ex.blurred(0), int(k0["x"]), int(k0["y"]), k0["angle"]), desc[0])


def test_flat_and_empty_images():
    ex = oracle.OrbExtractor(1000, 1.2, 8, 20, 7)
    kps, desc = ex(np.full((375, 1242), 9, np.uint8))
    assert len(kps) == 0
    with pytest.raises(ValueError):
        ex(np.zeros((40, 40), np.uint8))


# ------------------------------------------------------------------------------------------------ matching / BoW
def test_hamming_is_popcount():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    for x, y in zip(a, b):
        assert M.hamming256(x, y) == int(np.unpackbits(x ^ y).sum())


def test_bow_transform_and_score_properties():
    voc = M.Vocabulary.from_arrays(*M.random_vocabulary(10, 3, 0))
    rng = np.random.default_rng(2)
    d = rng.integers(0, 256, (800, 32), dtype=np.uint8)
    words, vals, fn, fo, fi = voc.transform(d, 2)
    assert (np.diff(words.astype(np.int64)) > 0).all() and vals.sum() == pytest.approx(1.0, abs=1e-12)
    assert (np.diff(fn.astype(np.int64)) > 0).all() and fo[-1] == len(fi) and len(set(fi.tolist())) == len(fi)
    w, wt, nid = voc.transform_features(d, 2)
    assert len(fi) == int((wt > 0).sum())  # stopped words (weight 0) are dropped (TemplatedVocabulary.h:1157)
    assert voc.score((words, vals), (words, vals)) == pytest.approx(1.0, abs=1e-12)
    other = voc.transform(rng.integers(0, 256, (800, 32), dtype=np.uint8), 2)
    s = voc.score((words, vals), other[:2])
    assert 0.0 <= s < 1.0 and s == voc.score(other[:2], (words, vals))


def test_search_by_bow_semantics():
    voc = M.Vocabulary.from_arrays(*M.random_vocabulary(10, 3, 0))
    rng = np.random.default_rng(3)
    d = rng.integers(0, 256, (600, 32), dtype=np.uint8)
    b = voc.transform(d, 2)
    ang = rng.uniform(0, 360, 600).astype(np.float32)
    A = M.Side(d, b[2], b[3], b[4], angles=ang)
    for variant in (0, 1, 2):
        m, n = M.search_by_bow(variant, A, A, 0.75, True)
        used = np.nonzero(b[1] >= 0)[0]
        assert n == int((m >= 0).sum()) and (m[m >= 0] == np.nonzero(m >= 0)[0]).all()  # self-match is the identity
    # ties: identical descriptors make best == second best, so the ratio test rejects (App. D.3)
    dd = np.tile(d[:1], (30, 1))
    bb = voc.transform(dd, 2)
    T = M.Side(dd, bb[2], bb[3], bb[4])
    assert M.search_by_bow(0, T, T, 0.9, False)[1] == 0
    # dead MapPoints on side A are skipped; variant 2 also honours side B
    valid = np.zeros(600, np.uint8)
    A0 = M.Side(d, b[2], b[3], b[4], valid=valid, angles=ang)
    assert M.search_by_bow(0, A0, A, 0.75, True)[1] == 0 and M.search_by_bow(2, A, A0, 0.75, True)[1] == 0


# ------------------------------------------------------------------------------------------------ bundle adjustment
def test_ba_jacobians_against_central_differences():
    oracle.lib()
    prob = ba_problem(10, 300, seed=1)
    rng = np.random.default_rng(0)
    for i in rng.integers(0, len(prob["edge_pose"]), 12):
        pi, li = prob["edge_pose"][i], prob["edge_point"][i]
        q, t, cam, X = prob["pose_q"][pi], prob["pose_t"][pi], prob["pose_cam"][pi], prob["point_xyz"][li]
        for obs, h, tol in ((np.array([*prob["edge_obs"][i][:2], -1.0]), 1e-6, 1e-5),       # monocular edge
                            (np.array([*prob["edge_obs"][i][:2], 300.0]), 1e-3, 5e-3)):     # stereo: invz is float32
            e, A, J = B.edge(q, t, cam, X, obs)
            An, Jn = np.zeros_like(A), np.zeros_like(J)
            for c in range(3):
                d = np.zeros(3); d[c] = h
                An[:, c] = (B.edge(q, t, cam, X + d, obs)[0] - B.edge(q, t, cam, X - d, obs)[0]) / (2 * h)
            for c in range(6):
                d = np.zeros(6); d[c] = h
                qp, tp = B.pose_oplus(q, t, d)
                qm, tm = B.pose_oplus(q, t, -d)
                Jn[:, c] = (B.edge(qp, tp, cam, X, obs)[0] - B.edge(qm, tm, cam, X, obs)[0]) / (2 * h)
            assert np.abs(A - An).max() < tol * max(1.0, np.abs(A).max())
            assert np.abs(J - Jn).max() < tol * max(1.0, np.abs(J).max())


def test_ba_converges_monotonically_and_respects_lm_constants():
    prob = ba_problem(20, 2000, seed=7, n_fusion=10)
    out, info = B.solve(prob, 10)
    chi = [info["chi2_initial"]] + [c for c, a in zip(info["trial_chi2"], info["trial_accepted"]) if a]
    assert all(b < a for a, b in zip(chi, chi[1:]))            # chi2 strictly decreases over accepted steps
    assert info["chi2_final"] == chi[-1] and info["iterations"] == 10
    assert B.chi2(out)[1] < 1.1 and B.chi2(prob)[1] > 5.0      # RMS reprojection error in pixels
    np.testing.assert_array_equal(out["pose_t"][0], prob["pose_t"][0])  # the fixed keyframe does not move
    assert prob["pose_t"] is not out["pose_t"]
    # lambda0 = 1e-5 * max diag(H) (optimization_algorithm_levenberg.cpp:47,166-180)
    assert 1e2 < info["lambda_initial"] < 1e5
    # stop flag polled before the first iteration
    out2, info2 = B.solve(prob, 10, stop=np.ones(1, np.uint8))
    assert info2["iterations"] == 0 and info2["stopped"] == 1


def test_ba_landmark_sharding_sums_to_the_same_system():
    """Two shards whose partial reduced systems are summed through the all-reduce hook equal the unsharded solve."""
    prob = ba_problem(15, 900, seed=4, n_fusion=6)
    full, info = B.solve(prob, 6)
    shards = [ba_shard(prob, r, 2) for r in range(2)]
    # emulate the collective in-process: run rank 0 and rank 1 in lock step with threads
    import threading
    bar = threading.Barrier(2)
    slots = [None, None]
    results = [None, None]

    def make_cb(rank):
        def cb(arr, op):
            slots[rank] = arr.copy()
            bar.wait()
            a, b = slots
            res = a + b if op == 0 else (np.minimum(a, b) if op == 1 else np.maximum(a, b))
            bar.wait()
            arr[:] = res
        return cb

    def run(rank):
        results[rank] = B.solve(shards[rank], 6, allreduce=make_cb(rank))

    ths = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for r in range(2):
        out, inf = results[r]
        assert inf["trial_accepted"] == info["trial_accepted"]
        np.testing.assert_allclose(out["pose_t"], full["pose_t"], atol=1e-8)
        np.testing.assert_allclose(out["point_xyz"], full["point_xyz"][shards[r]["_point_ids"]], atol=1e-8)
    np.testing.assert_array_equal(results[0][0]["pose_q"], results[1][0]["pose_q"])  # ranks end with identical poses


def test_projection_gemm_model_matches_cv2_golden(gold):
    """cv::Mat products in SearchByProjection (ORBmatcher.cc:1483,1488,1504) are cv::gemm on CV_32F: double accumulation,
    one rounding. Bit-exact against cv2.gemm 4.13.0 vectors."""
    from oracle import _proj_bind as PB
    R, t, x = gold["gemm_R"], gold["gemm_t"], gold["gemm_x"]
    for i in range(len(R)):
        M = np.concatenate([R[i], t[i]], 1)
        a = PB.gemm3(M, False, 1.0, x[i].ravel(), 1.0, t[i].ravel())
        np.testing.assert_array_equal(a.view(np.uint32), gold["gemm_Rx_plus_t"][i].ravel().view(np.uint32))
        b = PB.gemm3(M, True, -1.0, t[i].ravel())
        np.testing.assert_array_equal(b.view(np.uint32), gold["gemm_minus_Rt_t"][i].ravel().view(np.uint32))


def _projection_inputs(seed):
    from corb_slam_b200.frame import FrameView
    from corb_slam_b200.synth import projection_scene
    s = projection_scene(seed)
    c = s["cur"]
    fv = FrameView(c["x"], c["y"], c["octave"], c["angle"], c["desc"], c["u_right"], s["scales"], s["bounds"], s["K"], s["mbf"],
                   s["Tcw"], taken=s["taken"])
    return s, fv


def test_frame_grid_mirrors_assign_features_to_grid():
    """FrameView builds Frame::mGrid (Frame.cc:229-245, 386-397): every in-bounds feature sits in exactly one cell,
    indices ascend inside a cell, and GetFeaturesInArea over the grid equals a brute-force window query."""
    s, fv = _projection_inputs(3)
    assert fv.grid_off[0] == 0 and fv.grid_off[-1] == len(fv.grid_idx) <= fv.n
    assert len(np.unique(fv.grid_idx)) == len(fv.grid_idx)
    for c in np.nonzero(np.diff(fv.grid_off) > 1)[0][:200]:
        seg = fv.grid_idx[fv.grid_off[c]:fv.grid_off[c + 1]]
        assert np.all(np.diff(seg) > 0)
    ix, iy = 20, 17
    seg = fv.grid_idx[fv.grid_off[ix * 48 + iy]:fv.grid_off[ix * 48 + iy + 1]]
    px = np.floor((fv.x[seg] - fv.min_x) * fv.grid_w_inv + 0.5)
    assert np.all(px == ix)


def test_search_by_projection_semantics():
    """Oracle-level properties of both SearchByProjection variants on the synthetic scene: matches are injective where
    the MapPoints block, respect TH_HIGH, never touch features that were already taken, and mostly recover the true
    correspondences."""
    from oracle import _proj_bind as PB
    from oracle import _match_bind as M
    s, fv = _projection_inputs(5)
    cs = fv.c_struct()
    match, n = PB.search_by_projection_last(cs, fv.n, s["last_valid"], None, s["Xw"], s["mp_desc"], s["last"]["octave"],
                                            s["last"]["angle"], s["Tlw"], 15.0, False, True)
    got = np.nonzero(match >= 0)[0]
    assert n == len(got) > 800                      # all MapPoints block -> one feature per MapPoint, count consistent
    assert len(np.unique(match[got])) == len(got)
    assert not np.any(s["taken"][got])
    assert np.all(s["last_valid"][match[got]] == 1)
    d = np.array([M.hamming256(s["mp_desc"][match[i]], fv.desc[i]) for i in got[:300]])
    assert d.max() <= 100
    # without the orientation check there are at least as many matches
    _, n_no = PB.search_by_projection_last(cs, fv.n, s["last_valid"], None, s["Xw"], s["mp_desc"], s["last"]["octave"],
                                           s["last"]["angle"], s["Tlw"], 15.0, False, False)
    assert n_no >= n
    m2, n2 = PB.search_by_projection_map(cs, fv.n, s["in_view"], None, s["proj"], s["level"], s["view_cos"], s["mp_desc"], 1.0, 0.8)
    got2 = np.nonzero(m2 >= 0)[0]
    assert n2 == len(got2) > 800 and len(np.unique(m2[got2])) == len(got2)
    assert np.all(s["in_view"][m2[got2]] == 1) and not np.any(s["taken"][got2])
    # a larger search window (th = 3, Tracking.cc:1208) cannot lose candidates
    _, n3 = PB.search_by_projection_map(cs, fv.n, s["in_view"], None, s["proj"], s["level"], s["view_cos"], s["mp_desc"], 3.0, 0.8)
    assert n3 >= n2 - 5


def test_ba_minimum_equals_scipy_least_squares():
    """Independent pin of the BA port (g2o / Optimizer.cc cannot be compiled here: no Eigen3): the minimum the oracle's LM
    converges to is the minimum scipy.optimize.least_squares finds for a residual written independently in numpy from the
    reference's edge definitions (types_six_dof_expmap.cpp:103-234: u = fx x/z + cx, v = fy y/z + cy, ur = u - bf/z, weighted
    by invSigma2) - same objective, same local minimum from the same start, whatever path each optimiser takes."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation
    from corb_slam_b200.synth import ba_problem
    from oracle import _ba_bind as B
    oracle.lib()
    prob = ba_problem(8, 60, seed=5, obs_per_point=4, n_fusion=2)
    P, L = len(prob["pose_t"]), len(prob["point_xyz"])
    ep, el = prob["edge_pose"], prob["edge_point"]
    obs, w = prob["edge_obs"], np.sqrt(prob["edge_inv_sigma2"])
    cam = prob["pose_cam"]
    stereo = obs[:, 2] >= 0
    free_p = np.nonzero(prob["pose_fixed"] == 0)[0]
    rv0 = Rotation.from_quat(prob["pose_q"]).as_rotvec()   # (x, y, z, w), world -> camera

    def unpack(x):
        rv, t = rv0.copy(), prob["pose_t"].copy()
        rv[free_p] = x[:3 * len(free_p)].reshape(-1, 3)
        t[free_p] = x[3 * len(free_p):6 * len(free_p)].reshape(-1, 3)
        return rv, t, x[6 * len(free_p):].reshape(L, 3)

    def residual(x):
        rv, t, X = unpack(x)
        R = Rotation.from_rotvec(rv).as_matrix()
        Xc = np.einsum("eij,ej->ei", R[ep], X[el]) + t[ep]
        fx, fy, cx, cy, bf = (cam[ep, k] for k in range(5))
        u = fx * Xc[:, 0] / Xc[:, 2] + cx
        v = fy * Xc[:, 1] / Xc[:, 2] + cy
        r = [w * (obs[:, 0] - u), w * (obs[:, 1] - v), np.where(stereo, w * (obs[:, 2] - (u - bf / Xc[:, 2])), 0.0)]
        return np.concatenate(r)

    x0 = np.concatenate([rv0[free_p].ravel(), prob["pose_t"][free_p].ravel(), prob["point_xyz"].ravel()])
    chi2_0 = float((residual(x0) ** 2).sum())
    o0, _ = B.chi2(prob)
    # the same function up to the float32 reciprocal the reference keeps in the stereo projection (1.0f / z, :151): 5e-8 relative
    assert o0 == pytest.approx(chi2_0, rel=1e-6)
    sol = least_squares(residual, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=4000)
    o_out, o_info = B.solve(prob, 60, robust=False)
    o_min, _ = B.chi2(o_out)
    assert o_min < 0.5 * chi2_0
    assert o_min == pytest.approx(float((sol.fun ** 2).sum()), rel=1e-5)
    rv, t, X = unpack(sol.x)
    assert np.abs(o_out["pose_t"] - t).max() < 1e-4
    # (a weakly constrained landmark - two nearly parallel rays - sits in a flat valley: compare the bulk, not the worst one)
    assert np.percentile(np.abs(o_out["point_xyz"] - X).max(axis=1), 90) < 1e-3


def test_fast_pretest_of_the_cuda_kernel_is_a_superset_of_the_corners():
    """k_fast_cells_g rejects a pixel before the exact score unless (|r0-v| > th or |r8-v| > th) and (|r4-v| > th or |r12-v| > th)
    (DESIGN.md section 4.1). Every FAST-9/16 corner must pass that filter - checked here against the oracle's cv2-pinned
    detector (without NMS every positive score is a corner) on textured, noisy and synthetic bench images, at both thresholds."""
    from corb_slam_b200.synth import stereo_frame
    rng = np.random.default_rng(11)
    imgs = [stereo_frame(1234)[0][:200, :300], rng.integers(0, 256, (120, 160)).astype(np.uint8),
            (rng.integers(0, 2, (90, 130)) * 200 + rng.integers(0, 40, (90, 130))).astype(np.uint8)]
    n_corners = 0
    for img in imgs:
        I = img.astype(np.int32)
        v = I[3:-3, 3:-3]
        r0, r8, r4, r12 = I[6:, 3:-3], I[:-6, 3:-3], I[3:-3, 6:], I[3:-3, :-6]
        for th in (20, 7):
            far = lambda r: np.abs(r - v) > th
            keep = (far(r0) | far(r8)) & (far(r4) | far(r12))
            kps = oracle.fast_detect(img, th)           # (x, y, response) after NMS: a subset of the corners
            xs, ys = np.asarray([k[0] for k in kps], int), np.asarray([k[1] for k in kps], int)
            assert keep[ys - 3, xs - 3].all()
            # and every corner before NMS: response = (max over nine-arcs of the arc minimum) - 1, a corner at th iff response >= th
            sc = oracle.fast_score(img).astype(np.int32)[3:-3, 3:-3]
            assert keep[sc >= th].all()
            n_corners += int((sc >= th).sum())
    assert n_corners > 5000
