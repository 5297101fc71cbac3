"""The reference's own C++ running ON TOP of the drop-in (oracle/_ref/libshim.so = reference Frame.cc + rest of ORBmatcher.cc +
rest of PnPsolver.cc + DBoW2, hot-path bodies replaced by shim/*.cc, linked with libcorb_b200.so) against the untouched
reference (oracle/_ref/libref.so): the same C++ calls - Frame::Frame(imLeft, imRight, ...), ORBmatcher::SearchByBoW /
SearchByProjection, PnPsolver::iterate, Optimizer::BundleAdjustment - give the same results through the class seams."""
import numpy as np
import pytest

import oracle
from oracle import _ba_bind as B
from oracle import _match_bind as M
from oracle import ref, shim
from corb_slam_b200.frame import FrameView
from corb_slam_b200.synth import KITTI_CAM, pnp_problem, projection_scene, stereo_frame
from test_shim_cpu import _world

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (shim.available() and ref.available()), reason="oracle/_ref was not shipped")]
ORB = (2000, 1.2, 8, 20, 7)
S = shim.classes


@pytest.mark.parametrize("seed", [1234, 1237])
def test_frame_constructor_on_the_shim_extractor(seed):
    """Frame.cc (reference, unmodified) -> two std::threads -> shim ORBextractor::operator() -> GPU; then the reference's own
    ComputeStereoMatches reads the shim's mvImagePyramid (with its 19 px frame) and AssignFeaturesToGrid runs on the result."""
    fx, fy, cx, cy, bf = KITTI_CAM
    L, R = stereo_frame(seed)
    a = ref.Frame(ref.ORBextractor(*ORB), ref.ORBextractor(*ORB), L, R, fx, fy, cx, cy, bf)
    sl, sr = S.ORBextractor(*ORB), S.ORBextractor(*ORB)
    b = S.Frame(sl, sr, L, R, fx, fy, cx, cy, bf)
    assert a.keys.tobytes() == b.keys.tobytes() and a.keys_right.tobytes() == b.keys_right.tobytes() and len(a.keys) > 1900
    assert np.array_equal(a.desc, b.desc) and np.array_equal(a.desc_right, b.desc_right)
    assert a.u_right.tobytes() == b.u_right.tobytes() and a.depth.tobytes() == b.depth.tobytes() and (a.u_right >= 0).sum() > 1000
    assert np.array_equal(a.grid_off, b.grid_off) and np.array_equal(a.grid_idx, b.grid_idx)
    r = ref.ORBextractor(*ORB)
    r(L)
    for level in range(8):
        assert np.array_equal(sl.pyramid(level), r.pyramid(level))
    assert sl.scale.tobytes() == r.scale.tobytes() and sl.inv_sigma2.tobytes() == r.inv_sigma2.tobytes()  # the getters


def test_search_by_bow_members_on_the_gpu():
    voc = ref.ORBVocabulary(ref.vocabulary_text(stripped=True))
    ex = oracle.OrbExtractor(*ORB)
    base = stereo_frame(1234)[0]
    rng = np.random.default_rng(1)
    again = (np.roll(base, 5, axis=1).astype(np.int16) + rng.normal(0, 3.0, base.shape).round().astype(np.int16)).clip(0, 255).astype(np.uint8)
    frames = [ex(base), ex(again), ex(stereo_frame(1235)[0])]
    fvs = [voc.transform(d, 4)[2:] for _, d in frames]
    total = 0
    for variant in (0, 1, 2):
        for i, j in ((0, 1), (1, 0), (0, 2), (0, 0)):
            A = M.Side(frames[i][1], *fvs[i], valid=rng.random(len(frames[i][1])) < 0.7, angles=frames[i][0]["angle"])
            Bs = M.Side(frames[j][1], *fvs[j], valid=rng.random(len(frames[j][1])) < 0.7, angles=frames[j][0]["angle"])
            for nn, ori in ((0.7, True), (0.9, False)):
                m1, n1 = ref.search_by_bow(variant, A, Bs, nn, ori)
                m2, n2 = S.search_by_bow(variant, A, Bs, nn, ori)
                assert n1 == n2 and np.array_equal(m1, m2), (variant, i, j, nn, ori)
                total += n1
    assert total > 3000


def test_search_by_projection_members_on_the_gpu():
    for seed, th, mono, ori in ((1, 15.0, False, True), (3, 15.0, True, False), (4, 30.0, False, True)):
        s = projection_scene(seed)
        c = s["cur"]
        fv = FrameView(c["x"], c["y"], c["octave"], c["angle"], c["desc"], c["u_right"], s["scales"], s["bounds"], s["K"], s["mbf"], s["Tcw"],
                       taken=s["taken"])
        args = (fv.c_struct(), fv.n, s["last_valid"], s["last_blocks"], s["Xw"], s["mp_desc"], s["last"]["octave"], s["last"]["angle"],
                s["Tlw"], th, mono, ori)
        (rm, rn), (sm, sn) = ref.search_by_projection_last(*args), S.search_by_projection_last(*args)
        assert rn == sn > 300 and np.array_equal(rm, sm)
        args = (fv.c_struct(), fv.n, s["in_view"], None, s["proj"], s["level"], s["view_cos"], s["mp_desc"], 3.0, 0.8)
        (rm, rn), (sm, sn) = ref.search_by_projection_map(*args), S.search_by_projection_map(*args)
        assert rn == sn > 500 and np.array_equal(rm, sm)


def test_pnp_iterate_member_on_the_gpu():
    sig = ref.ORBextractor(*ORB).sigma2
    found = 0
    for c in range(12):
        n = [150, 60, 333, 31, 1000, 97][c % 6]
        p = pnp_problem(100 + c, n=n, outlier_fraction=[0.25, 0.45, 0.1, 0.97][c % 4], pixel_noise=0.5)
        octave = (np.arange(n) % 8).astype(np.int32)
        K = [float(v) for v in p["K"]]
        a, b = ref.PnPsolver(p["p2d"], octave, p["p3d"], sig, *K), S.PnPsolver(p["p2d"], octave, p["p3d"], sig, *K)
        assert a.SetRansacParameters(0.99, 10, 300, 4, 0.5, 5.991) == b.SetRansacParameters(0.99, 10, 300, 4, 0.5, 5.991)
        ra, rb = a.iterate(5, 1000 + c), b.iterate(5, 1000 + c)  # both consume the rand() stream from srand(seed)
        assert ra[0] == rb[0] and ra[1] == rb[1] and ra[3] == rb[3] and a.iterations == b.iterations and np.array_equal(ra[2], rb[2])
        if ra[0]:
            assert ra[4].tobytes() == rb[4].tobytes()
            found += 1
    assert found >= 6


def test_bundle_adjustment_member_on_the_gpu():
    """Optimizer::BundleAdjustment(vpKFs, vpMP, 10, NULL, nLoopKF, false) through shim/Optimizer_gba.cc on the GPU vs the oracle's
    solve of the same flattened graph: mTcwGBA / mPosGBA (float32) agree."""
    oracle.lib()
    prob, w, m = _world(P=60, L=6000, seed=9)
    f = w.flatten()
    flat = {k: f[k] for k in ("pose_q", "pose_t", "pose_fixed", "pose_cam", "point_xyz", "point_fixed", "edge_pose", "edge_point", "edge_obs",
                               "edge_inv_sigma2")}
    out, info = B.solve(flat, 10)
    w.run(10, 7, False)
    r = w.read()
    pos = {int(k): j for j, k in enumerate(m["kf_id"][m["perm"]])}
    worst = 0.0
    for j, kid in enumerate(f["pose_kf_id"]):
        i = pos[int(kid)]
        if m["kf_flags"][int(kid) - 1] & 1:
            assert r["kf_gba"][i] == 0
            continue
        assert r["kf_gba"][i] == 7
        worst = max(worst, float(np.abs(r["TcwGBA"][i] - shim.pose_from_quat(out["pose_q"][j], out["pose_t"][j])).max()))
    assert worst < 2e-5, worst
    pts = f["point_mp_index"]
    free = np.array([not (m["mp_flags"][l] & 1) for l in pts])
    assert np.abs(r["posGBA"][pts][free] - out["point_xyz"][free].astype(np.float32)).max() < 2e-4
    assert (r["mp_gba"][pts][free] == 7).all()
