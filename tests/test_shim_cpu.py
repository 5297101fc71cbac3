"""The C++ shims a maintainer adds to the reference tree (shim/*.cc), compiled against the reference's own
headers and linked with the reference's Frame.cc / rest of ORBmatcher.cc / rest of PnPsolver.cc / DBoW2 and libcorb_b200.so
into oracle/_ref/libshim.so. Without a GPU: the library loads, and the Optimizer::BundleAdjustment seam (SURVEY.md section 8
row a16: pointer graph -> flat arrays, float32 4x4 <-> SE3Quat, write-back of Optimizer.cc:216-263) round-trips through the
oracle's solver."""
import numpy as np
import pytest

import oracle
from oracle import _ba_bind as B
from oracle import shim
from corb_slam_b200.synth import ba_problem

pytestmark = pytest.mark.skipif(not shim.available(), reason="oracle/_ref/libshim.so is not built")


def _R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _T32(q, t):
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = _R(q).astype(np.float32)
    T[:3, 3] = np.asarray(t, np.float32)
    return T


def test_library_exports_the_class_seams():
    import subprocess
    out = subprocess.run(["nm", "-DC", shim.LIB_PATH], capture_output=True, text=True).stdout
    for sym in ("ORB_SLAM2::ORBextractor::operator()", "ORB_SLAM2::ORBextractor::ORBextractor(int, float, int, int, int)",
                "ORB_SLAM2::ORBmatcher::SearchByBoW(ORB_SLAM2::KeyFrame*, ORB_SLAM2::Frame&",
                "ORB_SLAM2::ORBmatcher::SearchByBoW(ORB_SLAM2::KeyFrame*, ORB_SLAM2::KeyFrame*",
                "ORB_SLAM2::ORBmatcher::SearchByBoWInServer(", "ORB_SLAM2::ORBmatcher::SearchByProjection(ORB_SLAM2::Frame&, ORB_SLAM2::Frame const&",
                "ORB_SLAM2::ORBmatcher::SearchByProjection(ORB_SLAM2::Frame&, std::vector<ORB_SLAM2::MapPoint*",
                "ORB_SLAM2::ORBmatcher::Fuse(",  # the rest of ORBmatcher.cc is the reference's own code
                "ORB_SLAM2::PnPsolver::iterate(", "ORB_SLAM2::PnPsolver::compute_pose(", "ORB_SLAM2::Frame::ComputeStereoMatches()",
                "ORB_SLAM2::Optimizer::BundleAdjustment("):
        assert sym in out, sym
    assert "corb_orb_extract" in out and "corb_ba_solve" in out and "corb_bow_match" in out  # undefined: resolved by libcorb_b200.so
    needed = subprocess.run(["readelf", "-d", shim.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcorb_b200.so" in needed


def test_se3quat_conversions_follow_eigen_and_g2o():
    rng = np.random.default_rng(0)
    for i in range(300):
        q = rng.normal(size=4)
        if i % 5 == 0:
            q[3] = abs(q[3]) * 1e-3  # rotations near 180 degrees: the trace <= 0 branches
        q /= np.linalg.norm(q)
        if q[3] < 0:
            q = -q
        t = rng.normal(size=3) * 10
        T = _T32(q, t)
        q2, t2 = shim.quat_from_pose(T)
        assert abs(np.linalg.norm(q2) - 1) < 1e-15 and q2[3] >= 0                 # normalizeRotation (se3quat.h:280-288)
        assert np.abs(_R(q2) - T[:3, :3].astype(np.float64)).max() < 2e-7          # float32 rotation matrices are not exactly orthogonal
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 2e-7
        assert np.array_equal(t2, T[:3, 3].astype(np.float64))                    # translation: float widened, untouched
        T2 = shim.pose_from_quat(q2, t2)
        assert np.array_equal(T2[:3, :3], _R(q2).astype(np.float32)) and np.array_equal(T2[3], [0, 0, 0, 1])
        assert np.abs(T2 - T).max() < 2e-6


def _world(P=30, L=1500, seed=5):
    prob = ba_problem(P, L, seed=seed, n_fusion=6)
    rng = np.random.default_rng(seed)
    kf_id = np.arange(1, P + 1, dtype=np.uint64)      # mnId 1 is the fixed anchor (Optimizer.cc:92)
    perm = rng.permutation(P)                          # vpKFs arrives in arbitrary order
    kf_T = np.stack([_T32(prob["pose_q"][i], prob["pose_t"][i]) for i in range(P)])
    kf_flags = np.zeros(P, np.uint8)
    kf_flags[7] = 2   # a bad keyframe: no vertex, its observations are skipped (:87, :130)
    kf_flags[11] = 1  # a keyframe fixed by getFixed() (client-side GBA, Cache.cc:482)
    mp_flags = np.zeros(L, np.uint8)
    mp_flags[5] = 4   # a NULL entry of vpMP
    mp_flags[9] = 2   # a bad map point
    mp_flags[13] = 1  # a fixed map point
    lonely = int(np.setdiff1d(np.arange(L), prob["edge_point"])[0]) if len(np.setdiff1d(np.arange(L), prob["edge_point"])) else None
    octave = rng.integers(0, 8, len(prob["edge_pose"])).astype(np.int32)
    inv_sigma2 = (1.0 / (np.float32(1.2) ** (2 * np.arange(8, dtype=np.float32)))).astype(np.float32)
    uvr = prob["edge_obs"].astype(np.float32)
    uvr[np.isnan(uvr[:, 2]) | (prob["edge_obs"][:, 2] < 0), 2] = -1.0
    w = shim.World(kf_id[perm], kf_T[perm], kf_flags[perm], prob["pose_cam"].astype(np.float32)[perm], prob["point_xyz"].astype(np.float32),
                   mp_flags, np.argsort(perm)[prob["edge_pose"]], prob["edge_point"], uvr, octave, inv_sigma2)
    return prob, w, dict(kf_id=kf_id, perm=perm, kf_T=kf_T, kf_flags=kf_flags, mp_flags=mp_flags, octave=octave, inv_sigma2=inv_sigma2, uvr=uvr)


def test_ba_flatten_solve_write_back_round_trip():
    oracle.lib()
    prob, w, m = _world()
    P, L = len(m["kf_id"]), len(m["mp_flags"])
    f = w.flatten()
    # ---- poses: live keyframes in ascending mnId, float32 4x4 -> unit quaternion + translation, fixed rule, intrinsics
    live = [i for i in range(P) if not (m["kf_flags"][i] & 2)]
    assert f["pose_kf_id"].tolist() == [int(m["kf_id"][i]) for i in live]
    for j, i in enumerate(live):
        q, t = shim.quat_from_pose(m["kf_T"][i])
        assert np.array_equal(f["pose_q"][j], q) and np.array_equal(f["pose_t"][j], t)
        assert f["pose_fixed"][j] == (1 if (m["kf_id"][i] == 1 or m["kf_flags"][i] & 1) else 0)
        assert np.array_equal(f["pose_cam"][j], prob["pose_cam"][i].astype(np.float32).astype(np.float64))
    # ---- points and observations: NULL / bad points and points without a usable observation are left out
    dense = {int(k): j for j, k in enumerate(f["pose_kf_id"])}
    want_edges, want_points = [], []
    by_point = {}
    for e in range(len(prob["edge_pose"])):
        by_point.setdefault(int(prob["edge_point"][e]), []).append(e)
    for l in range(L):
        if m["mp_flags"][l] & 6:
            continue
        es = [e for e in by_point.get(l, []) if int(m["kf_id"][prob["edge_pose"][e]]) in dense]
        if not es:
            continue
        want_points.append(l)
        for e in es:
            want_edges.append((dense[int(m["kf_id"][prob["edge_pose"][e]])], len(want_points) - 1, e))
    assert f["point_mp_index"].tolist() == want_points
    assert np.array_equal(f["point_xyz"], prob["point_xyz"].astype(np.float32).astype(np.float64)[want_points])
    assert f["point_fixed"].tolist() == [int(m["mp_flags"][l] & 1) for l in want_points]
    # std::map<KeyFrame*, size_t> walks a point's observations in heap-address order: compare as sets per point
    got = sorted(zip(f["edge_point"].tolist(), f["edge_pose"].tolist(), map(tuple, f["edge_obs"].tolist()), f["edge_inv_sigma2"].tolist()))
    exp = sorted((pt, ps, tuple(m["uvr"][e].astype(np.float64).tolist()), float(m["inv_sigma2"][m["octave"][e]])) for ps, pt, e in want_edges)
    assert got == exp and len(got) > 5000
    assert np.all(np.diff(f["edge_point"]) >= 0)  # grouped by landmark: corb_ba_solve uses the arrays in place
    # ---- solve on the flat arrays with the oracle (the GPU test runs corb_ba_solve through the same seam), then write back
    flat = {k: f[k] for k in ("pose_q", "pose_t", "pose_fixed", "pose_cam", "point_xyz", "point_fixed", "edge_pose", "edge_point",
                               "edge_inv_sigma2")}
    flat["edge_obs"] = np.where(f["edge_obs"] < 0, np.nan, f["edge_obs"]) if np.isnan(prob["edge_obs"]).any() else f["edge_obs"]
    out, info = B.solve(flat, 5)
    assert info["chi2_final"] < 0.05 * info["chi2_initial"]
    before = w.read()
    w.write_back(out["pose_q"], out["pose_t"], out["point_xyz"], nLoopKF=42)     # loop-closing flavour: mTcwGBA / mPosGBA (:233-237, :257-261)
    r = w.read()
    assert np.array_equal(r["Tcw"], before["Tcw"]) and np.array_equal(r["xyz"], before["xyz"]) and r["cache_kfs"] == 0
    pos = {int(k): j for j, k in enumerate(m["kf_id"][m["perm"]])}
    for j, kid in enumerate(f["pose_kf_id"]):
        i = pos[int(kid)]
        fixed = m["kf_flags"][int(kid) - 1] & 1
        if fixed:
            assert r["kf_gba"][i] == 0 and not r["TcwGBA"][i].any()               # getFixed(): skipped (:222)
        else:
            assert r["kf_gba"][i] == 42 and np.array_equal(r["TcwGBA"][i], shim.pose_from_quat(out["pose_q"][j], out["pose_t"][j]))
    for j, l in enumerate(want_points):
        if m["mp_flags"][l] & 1:
            assert r["mp_gba"][l] == 0
        else:
            assert r["mp_gba"][l] == 42 and np.array_equal(r["posGBA"][l], out["point_xyz"][j].astype(np.float32))
    assert r["mp_gba"][5] == 0 and r["mp_gba"][9] == 0
    w.write_back(out["pose_q"], out["pose_t"], out["point_xyz"], nLoopKF=0)      # map-fusion flavour: SetPose / SetWorldPos + cache queues
    r = w.read()
    n_free_kf = sum(1 for kid in f["pose_kf_id"] if not (m["kf_flags"][int(kid) - 1] & 1))
    n_free_mp = sum(1 for l in want_points if not (m["mp_flags"][l] & 1))
    assert r["cache_kfs"] == n_free_kf and r["cache_mps"] == n_free_mp and r["normal_updates"].sum() == n_free_mp
    j = 3
    assert np.array_equal(r["Tcw"][pos[int(f["pose_kf_id"][j])]], shim.pose_from_quat(out["pose_q"][j], out["pose_t"][j]))
