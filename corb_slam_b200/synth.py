"""Seeded synthetic inputs for the benchmark and parity tests. numpy only, no cv2."""
import numpy as np


def stereo_frame(seed, w=1242, h=375, disparity=8):
    """One synthetic stereo pair: 1500 random grey rectangles on 128, 3x3 Gaussian (sigma 0.8), N(0, 2^2) noise.

    Right image = left image shifted by `disparity` px (np.roll), as SURVEY.md §8d specifies.
    """
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128.0, np.float64)
    n = 1500
    xs = rng.integers(0, w, n)
    ys = rng.integers(0, h, n)
    ws = rng.integers(6, 90, n)
    hs = rng.integers(6, 60, n)
    gs = rng.integers(0, 256, n)
    for x, y, rw, rh, g in zip(xs, ys, ws, hs, gs):
        img[y:y + rh, x:x + rw] = g
    k = np.exp(-0.5 * (np.array([-1.0, 0.0, 1.0]) / 0.8) ** 2)
    k /= k.sum()
    p = np.pad(img, 1, mode="reflect")
    img = k[0] * p[1:-1, :-2] + k[1] * p[1:-1, 1:-1] + k[2] * p[1:-1, 2:]
    p = np.pad(img, 1, mode="reflect")
    img = k[0] * p[:-2, 1:-1] + k[1] * p[1:-1, 1:-1] + k[2] * p[2:, 1:-1]
    img = img + rng.normal(0.0, 2.0, img.shape)
    left = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    right = np.roll(left, -disparity, axis=1)
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


def frame_seed(idx):
    return 1234 + idx


# --------------------------------------------------------------------------------------------- global BA problem
KITTI_CAM = (718.856, 718.856, 607.1928, 185.2157, 386.1448)  # fx fy cx cy bf  (KITTI00-02.yaml:8-11,25)


def _quat_from_yaw_pitch(yaw, pitch):
    """(x,y,z,w) of R = Ry(yaw) * Rx(pitch)."""
    cy, sy, cp, sp = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2)
    # q_y = (0, sy, 0, cy), q_x = (sp, 0, 0, cp); q = q_y * q_x
    return np.stack([cy * sp, sy * cp, -sy * sp, cy * cp], -1)


def _quat_to_R(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def ba_problem(n_poses=2000, n_points=200000, seed=7, obs_per_point=5, n_fusion=40, stereo_fraction=0.8, perturb=True):
    """Synthetic global-BA problem of SURVEY.md §8d: a 1 m/keyframe "street" trajectory, landmarks seen by
    `obs_per_point` consecutive keyframes, KITTI intrinsics, 1 px observation noise, octave-dependent information,
    `n_fusion` long-range observations (map-fusion / loop-closure links), first keyframe fixed.

    Returns a dict of numpy arrays in the layout of corb_ba_problem (include/corb_b200.h)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, bf = KITTI_CAM
    P, L = n_poses, n_points
    # ground-truth poses: camera looks along +Z of the world, small yaw/pitch wobble
    yaw = rng.normal(0, np.deg2rad(1.0), P)
    pitch = rng.normal(0, np.deg2rad(0.3), P)
    q_wc = _quat_from_yaw_pitch(yaw, pitch)               # camera -> world
    C = np.stack([rng.normal(0, 0.05, P), rng.normal(0, 0.02, P), np.arange(P) * 1.0], -1)
    R_wc = _quat_to_R(q_wc)
    q_cw = q_wc * np.array([-1.0, -1.0, -1.0, 1.0])        # world -> camera (conjugate)
    R_cw = np.transpose(R_wc, (0, 2, 1))
    t_cw = -np.einsum("pij,pj->pi", R_cw, C)
    # landmarks: pixel + depth in the nearest observer, back-projected
    k = rng.integers(0, P, L)
    u = rng.uniform(20, 1221, L); v = rng.uniform(20, 356, L); d = rng.uniform(5.0, 30.0, L)
    Xc = np.stack([(u - cx) * d / fx, (v - cy) * d / fy, d], -1)
    Xw = np.einsum("lij,lj->li", R_wc[k], Xc) + C[k]
    ep, el = [], []
    for j in range(obs_per_point):
        pi = k - j
        ok = pi >= 0
        ep.append(pi[ok]); el.append(np.nonzero(ok)[0])
    if n_fusion > 0:  # long-range links make the reduced camera system non-banded
        lm = rng.choice(L, size=min(n_fusion, L), replace=False)
        far = k[lm] - rng.integers(max(P // 4, obs_per_point + 1), max(P // 2, obs_per_point + 2), len(lm))
        ok = far >= 0
        ep.append(far[ok]); el.append(lm[ok])
    edge_pose = np.concatenate(ep).astype(np.int32)
    edge_point = np.concatenate(el).astype(np.int32)
    order = np.lexsort((edge_pose, edge_point))            # MapPoint-major like Optimizer.cc:106-197
    edge_pose, edge_point = edge_pose[order], edge_point[order]
    E = len(edge_pose)
    Xc_e = np.einsum("eij,ej->ei", R_cw[edge_pose], Xw[edge_point]) + t_cw[edge_pose]
    z = Xc_e[:, 2]
    assert (z > 1.0).all()
    pu = fx * Xc_e[:, 0] / z + cx + rng.normal(0, 1.0, E)
    pv = fy * Xc_e[:, 1] / z + cy + rng.normal(0, 1.0, E)
    pr = pu - bf / z + rng.normal(0, 1.0, E)
    mono = rng.random(E) >= stereo_fraction
    pr[mono] = -1.0
    # measurements are float32 in the reference (cv::KeyPoint, mvuRight)
    obs = np.stack([pu, pv, pr], -1).astype(np.float32).astype(np.float64)
    octave = rng.integers(0, 8, E)
    sigma2 = np.cumprod(np.concatenate([[1.0], np.full(7, 1.2)])).astype(np.float32) ** 2
    inv_sigma2 = (np.float32(1.0) / sigma2.astype(np.float32))[octave].astype(np.float64)
    pose_q, pose_t, pts = q_cw.copy(), t_cw.copy(), Xw.copy()
    if perturb:
        dC = rng.normal(0, 0.05, (P, 3))
        dyaw = rng.normal(0, np.deg2rad(0.5), P); dpitch = rng.normal(0, np.deg2rad(0.5), P)
        q_wc_n = _quat_from_yaw_pitch(yaw + dyaw, pitch + dpitch)
        Cn = C + dC
        q_wc_n[0], Cn[0] = q_wc[0], C[0]                   # the fixed keyframe keeps its pose
        R_cw_n = np.transpose(_quat_to_R(q_wc_n), (0, 2, 1))
        pose_q = q_wc_n * np.array([-1.0, -1.0, -1.0, 1.0])
        pose_t = -np.einsum("pij,pj->pi", R_cw_n, Cn)
        pts = Xw + rng.normal(0, 0.10, (L, 3))
    # poses/points enter the reference as float32 cv::Mat (Converter.cc:37-47)
    pose_t = pose_t.astype(np.float32).astype(np.float64)
    pts = pts.astype(np.float32).astype(np.float64)
    pose_fixed = np.zeros(P, np.uint8); pose_fixed[0] = 1      # mnId == 1 (Optimizer.cc:94); index 0 here
    return {
        "pose_q": np.ascontiguousarray(pose_q), "pose_t": np.ascontiguousarray(pose_t), "pose_fixed": pose_fixed,
        "pose_cam": np.tile(np.array(KITTI_CAM, np.float64), (P, 1)), "point_xyz": np.ascontiguousarray(pts),
        "point_fixed": np.zeros(L, np.uint8), "edge_pose": edge_pose, "edge_point": edge_point,
        "edge_obs": np.ascontiguousarray(obs), "edge_inv_sigma2": np.ascontiguousarray(inv_sigma2),
    }


def ba_shard(prob, rank, world):
    """Landmark sharding of SURVEY.md §8e: landmark l goes to rank l % world with its edges; poses are replicated."""
    keep_pts = np.nonzero(np.arange(len(prob["point_xyz"])) % world == rank)[0]
    remap = np.full(len(prob["point_xyz"]), -1, np.int64)
    remap[keep_pts] = np.arange(len(keep_pts))
    ke = remap[prob["edge_point"]] >= 0
    out = dict(prob)
    out["point_xyz"] = np.ascontiguousarray(prob["point_xyz"][keep_pts])
    out["point_fixed"] = np.ascontiguousarray(prob["point_fixed"][keep_pts])
    out["edge_pose"] = np.ascontiguousarray(prob["edge_pose"][ke])
    out["edge_point"] = np.ascontiguousarray(remap[prob["edge_point"][ke]].astype(np.int32))
    out["edge_obs"] = np.ascontiguousarray(prob["edge_obs"][ke])
    out["edge_inv_sigma2"] = np.ascontiguousarray(prob["edge_inv_sigma2"][ke])
    out["pose_q"] = prob["pose_q"].copy(); out["pose_t"] = prob["pose_t"].copy()
    out["_point_ids"] = keep_pts
    return out


# --------------------------------------------------------------------------------------------- projection matching scene
def projection_scene(seed=0, n_points=2500, w=1242, h=375, n_levels=8, scale_factor=1.2, motion=0.6, desc_noise=18,
                     clutter=600, temporal_fraction=0.15):
    """Two consecutive stereo frames of a synthetic scene for ORBmatcher::SearchByProjection: random 3-D points in front
    of a KITTI-like camera, a last frame (identity pose) and a current frame moved `motion` metres forward with a small
    rotation. Every point yields a last-frame feature with a MapPoint; the current frame holds noisy re-detections of
    most points plus `clutter` unrelated features. Returns a dict of numpy arrays (all float32 / int32 / uint8)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, bf = KITTI_CAM
    scales = np.float32(scale_factor) ** np.arange(n_levels, dtype=np.float32)
    # world points (last camera frame = world)
    z = rng.uniform(4.0, 60.0, n_points)
    x = (rng.uniform(0, w, n_points) - cx) / fx * z
    y = (rng.uniform(0, h, n_points) - cy) / fy * z
    Xw = np.stack([x, y, z], 1).astype(np.float32)
    Tlw = np.eye(4, dtype=np.float32)[:3]
    yaw = 0.01
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    Tcw = np.concatenate([R, np.array([[0.02], [0.01], [-motion]])], 1).astype(np.float32)

    def project(T, X):
        Xc = X @ T[:, :3].T.astype(np.float64) + T[:, 3].astype(np.float64)
        return fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy, Xc[:, 2]

    desc = rng.integers(0, 256, (n_points, 32), dtype=np.uint8)

    def noisy(d, nbits):
        d = d.copy()
        flips = rng.integers(0, 256, (len(d), nbits))
        for j in range(nbits):
            d[np.arange(len(d)), flips[:, j] >> 3] ^= (1 << (flips[:, j] & 7)).astype(np.uint8)
        return d

    octave = np.clip(np.floor(np.log(np.maximum(60.0 / z, 1.0)) / np.log(scale_factor)), 0, n_levels - 1).astype(np.int32)
    ul, vl, zl = project(Tlw, Xw)
    last = {"x": ul.astype(np.float32), "y": vl.astype(np.float32), "octave": octave,
            "angle": rng.uniform(0, 360, n_points).astype(np.float32)}
    last_valid = (rng.random(n_points) < 0.9).astype(np.uint8)             # pMP && !outlier
    last_blocks = (rng.random(n_points) >= temporal_fraction).astype(np.uint8)  # Observations() > 0 (temporal points: 0)
    # current frame: re-detections (80 %) + clutter, shuffled
    uc, vc, zc = project(Tcw, Xw)
    seen = (rng.random(n_points) < 0.8) & (uc > 0) & (uc < w) & (vc > 0) & (vc < h) & (zc > 0.5)
    ids = np.nonzero(seen)[0]
    cx_ = np.concatenate([uc[ids] + rng.normal(0, 1.5, len(ids)), rng.uniform(0, w, clutter)])
    cy_ = np.concatenate([vc[ids] + rng.normal(0, 1.5, len(ids)), rng.uniform(0, h, clutter)])
    coct = np.concatenate([np.clip(octave[ids] + rng.integers(-1, 2, len(ids)), 0, n_levels - 1), rng.integers(0, n_levels, clutter)])
    cang = np.concatenate([last["angle"][ids] + rng.normal(0, 4, len(ids)), rng.uniform(0, 360, clutter)]) % 360.0
    cdesc = np.concatenate([noisy(desc[ids], desc_noise), rng.integers(0, 256, (clutter, 32), dtype=np.uint8)])
    depth = np.concatenate([zc[ids], rng.uniform(4, 60, clutter)])
    cur_ur = np.where(rng.random(len(cx_)) < 0.8, cx_ - bf / depth + rng.normal(0, 0.5, len(cx_)), -1.0)
    perm = rng.permutation(len(cx_))
    cur = {"x": cx_[perm].astype(np.float32), "y": cy_[perm].astype(np.float32), "octave": coct[perm].astype(np.int32),
           "angle": cang[perm].astype(np.float32), "desc": np.ascontiguousarray(cdesc[perm]),
           "u_right": cur_ur[perm].astype(np.float32)}
    taken = (rng.random(len(cx_)) < 0.03).astype(np.uint8)
    # local-map variant: Frame::isInFrustum outputs for the same points (ProjX/Y/XR, predicted level, viewing cosine)
    proj = np.stack([uc, vc, uc - bf / np.maximum(zc, 0.1)], 1).astype(np.float32)
    in_view = (seen | (rng.random(n_points) < 0.05)).astype(np.uint8) & (zc > 0.5)
    level = np.clip(octave + rng.integers(0, 2, n_points), 0, n_levels - 1).astype(np.int32)
    view_cos = rng.uniform(0.99, 1.0, n_points).astype(np.float32)
    return {"scales": scales, "bounds": (0.0, 0.0, float(w), float(h)), "K": (fx, fy, cx, cy), "mbf": bf, "Tcw": Tcw, "Tlw": Tlw,
            "cur": cur, "taken": taken, "last": last, "last_valid": last_valid, "last_blocks": last_blocks, "Xw": Xw,
            "mp_desc": desc, "proj": proj, "in_view": in_view.astype(np.uint8), "level": level, "view_cos": view_cos}


def pnp_problem(seed, n=120, outlier_fraction=0.4, pixel_noise=0.7, w=1242, h=375, n_levels=8, scale_factor=1.2):
    """One relocalisation / map-fusion candidate for PnPsolver (PnPsolver.cc:66-111): `n` 2-D keypoints matched to
    MapPoints, `outlier_fraction` of them wrong matches. Returns float32 p2d [n,2], p3d [n,3] (world), sigma2 [n]
    (mvLevelSigma2 of the keypoint octave), the KITTI intrinsics (fx, fy, cx, cy) and the true Tcw (3x4, float64)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, _ = KITTI_CAM
    yaw, pitch = rng.normal(0, 0.2), rng.normal(0, 0.03)
    Ry = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(pitch), -np.sin(pitch)], [0, np.sin(pitch), np.cos(pitch)]])
    R = Rx @ Ry
    t = rng.normal(0, 1.0, 3) * np.array([1.0, 0.1, 1.0])
    z = rng.uniform(4.0, 50.0, n)
    u = rng.uniform(0, w, n)
    v = rng.uniform(0, h, n)
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    Xw = (Xc - t) @ R  # Xc = R Xw + t
    octave = rng.integers(0, n_levels, n)
    sigma2 = (np.float32(scale_factor) ** octave.astype(np.float32)) ** 2
    p2d = np.stack([u, v], 1) + rng.normal(0, pixel_noise, (n, 2)) * np.sqrt(sigma2)[:, None]
    bad = rng.random(n) < outlier_fraction
    p2d[bad] = np.stack([rng.uniform(0, w, bad.sum()), rng.uniform(0, h, bad.sum())], 1)
    return {"p2d": p2d.astype(np.float32), "p3d": Xw.astype(np.float32), "sigma2": sigma2.astype(np.float32),
            "K": (np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy)), "Tcw": np.concatenate([R, t[:, None]], 1),
            "outlier": bad}
