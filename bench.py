#!/usr/bin/env python
"""Benchmark of the CORB-SLAM hot path on B200: synthetic 1242x375 stereo frames/s (BASELINE.json metric, config #2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun for N > 1); every rank is one `corbslam_client` stream pinned to its GPU (replicas:
frames of different robots are independent), so scaling is weak and there is no data-path collective.
A step = one stereo frame through ORBextractor::operator() for the left and the right image (Frame.cc:78-81).

  value : frames/s with both images already resident in HBM and results left in HBM (CUDA-event timed per step,
          L2 flushed between steps, max over ranks)
  e2e   : frames/s through the reference-facing C-ABI call with HOST buffers (H2D image, D2H keypoints+descriptors
          inside the timed region, wall clock)
  --impl reference : the CPU oracle port of the reference path (the reference itself cannot be built here: no
          OpenCV/Eigen/ROS), left and right image on two threads per client like Frame.cc:78-81.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1242, 375
ORB_PARAMS = (2000, 1.2, 8, 20, 7)  # KITTI00-02.yaml:38-51
N_POOL = 8  # distinct synthetic frames cycled through
ALGO_BYTES_PER_IMAGE = W * H + 1441432 + 60 * 2000  # input + pyramid + 60 B per keypoint (K = 2000)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    def __init__(self, index):
        self.rows = []
        self.proc = None
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel_label):
    """DRAM bytes per launch of `kernel_label` ("k_octtree[3]") from the committed ncu --set full capture of one stereo
    frame (profiles/r01b_stereo_frame_ncu_full.csv; every launch there covers both images of the pair, so the figure is
    halved to match the per-image algorithmic bytes). None when the capture is missing."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01b_stereo_frame_ncu_full.csv")
    if not os.path.exists(path):
        return None
    base = kernel_label.split("[")[0]
    idx = int(kernel_label.split("[")[1].rstrip("]")) if "[" in kernel_label else 0
    if base == "k_resize":
        idx -= 1  # levels 1..7
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    rd = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum")]
    wr = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum")]
    if not rd or not wr:
        return None

    def to_bytes(v, h):
        unit = h[h.index("[") + 1:h.index("]")] if "[" in h else "byte"
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    hits = [r for r in rows[1:] if r and r[0].startswith(base)]
    if idx >= len(hits):
        return None
    r = hits[idx]
    return 0.5 * (to_bytes(r[rd[0]], hdr[rd[0]]) + to_bytes(r[wr[0]], hdr[wr[0]]))


def cpu_extract_fps(n_frames, clients=1, stereo=False):
    """Oracle port of the reference path: `clients` concurrent clients, each running left/right on two threads
    (Frame.cc:78-81); with stereo=True also Frame::ComputeStereoMatches (Frame.cc:90) on the calling thread."""
    import oracle
    from corb_slam_b200.synth import stereo_frame, frame_seed
    frames = [stereo_frame(frame_seed(i)) for i in range(min(n_frames, N_POOL))]
    exs = [(oracle.OrbExtractor(*ORB_PARAMS), oracle.OrbExtractor(*ORB_PARAMS)) for _ in range(clients)]
    for exl, exr in exs:
        exl(frames[0][0]); exr(frames[0][1])

    def client(exl, exr):
        for i in range(n_frames):
            l, r = frames[i % len(frames)]
            res = {}
            t = threading.Thread(target=lambda: res.__setitem__("r", exr(r)))
            t.start()
            kl, dl = exl(l)
            t.join()
            if stereo:
                kr, dr = res["r"]
                oracle.stereo_matches(exl, exr, kl, dl, kr, dr, 386.1448, 386.1448 / 718.856)

    ths = [threading.Thread(target=client, args=e) for e in exs]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return clients * n_frames / dt, dt


BA_P, BA_L, BA_ITERS = 2000, 200000, 10  # SURVEY.md §8d / BASELINE.json config #5 synthetic global-BA problem


def ba_algorithmic_bytes(n_edges, n_points, n_poses, nnz_blocks):
    """bytes one LM iteration must move (edges, landmarks, poses, reduced-system blocks)."""
    return 552 * n_edges + 216 * n_points + 440 * n_poses + 288 * nnz_blocks


def cpu_ba(n_poses, n_points):
    """Oracle port of Optimizer::BundleAdjustment (single thread like g2o, Thirdparty/g2o/config.h:4)."""
    import oracle
    from oracle import _ba_bind as B
    from corb_slam_b200.synth import ba_problem
    oracle.lib()
    prob = ba_problem(n_poses, n_points, seed=7)
    t0 = time.perf_counter()
    out, info = B.solve(prob, BA_ITERS, robust=False)
    dt = time.perf_counter() - t0
    return {"value": 1e3 * dt / (n_points / 1e4), "unit": "ms/10k landmarks", "cores": 1, "kind": "port",
            "sample": "P=%d keyframes, L=%d landmarks, E=%d observations, %d LM iterations, %.1f s"
                      % (n_poses, n_points, len(prob["edge_pose"]), info["iterations"], dt),
            "iterations": info["iterations"], "chi2_final": info["chi2_final"], "rms_px": B.chi2(out)[1]}


def matcher_workload(seed=3, n_feat=2000, n_cand=32):
    """Synthetic matching/BoW workload: a k=10, L=5 vocabulary (111 110 nodes), one query frame of
    2000 descriptors and `n_cand` candidate keyframes that are noisy re-observations of it."""
    import oracle
    from oracle import _match_bind as M
    oracle.lib()
    spec = M.random_vocabulary(10, 5, seed)
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (n_feat, 32), dtype=np.uint8)
    cands = []
    for c in range(n_cand):
        d = base[rng.permutation(n_feat)].copy()
        flips = rng.integers(0, 256, (n_feat, 6))
        for j in range(6):
            d[np.arange(n_feat), flips[:, j] >> 3] ^= (1 << (flips[:, j] & 7)).astype(np.uint8)
        cands.append(d)
    return spec, base, cands


def bench_matcher(device, with_cpu=True):
    """Calls/s and Hamming pairs/s of SearchByBoW (batched over candidates), descriptors/s of the vocabulary descent,
    scores/s of the L1 BoW score; each beside the oracle port on one host thread."""
    import oracle
    from oracle import _match_bind as M
    from corb_slam_b200 import BowFeatures, ORBmatcher, ORBVocabulary
    spec, base, cands = matcher_workload()
    gvoc = ORBVocabulary.from_arrays(*spec, device=device)
    ovoc = M.Vocabulary.from_arrays(*spec)
    levelsup = 3  # L - 3 = level 2 nodes -> 100 groups, like ORBvoc (L = 6, levelsup = 4)
    n = len(base)
    rng = np.random.default_rng(0)
    ang = rng.uniform(0, 360, n).astype(np.float32)
    valid = (rng.random(n) < 0.7).astype(np.uint8)
    out = {}
    # ---- transform (voc->transform, Frame.cc:404)
    gvoc.transform(base, levelsup)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        qb = gvoc.transform(base, levelsup)
    dt = (time.perf_counter() - t0) / reps
    out["transform"] = {"descriptors_per_s": n / dt, "ms_per_call": dt * 1e3}
    cb = [gvoc.transform(c, levelsup) for c in cands]
    # ---- SearchByBoW, batched over candidates (MapFusion.cpp:691 / Tracking.cc:1405)
    m = ORBmatcher(0.75, True, device=device)
    A = [BowFeatures(c, b[2], b[3], b[4], valid=valid, angles=ang) for c, b in zip(cands, cb)]
    B = [BowFeatures(base, qb[2], qb[3], qb[4], angles=ang) for _ in cands]
    pairs = 0
    for a_, b_ in zip(cb, [qb] * len(cb)):  # Hamming evaluations = sum over common nodes of nA * nB
        na = dict(zip(a_[2].tolist(), np.diff(a_[3]).tolist())); nb = dict(zip(b_[2].tolist(), np.diff(b_[3]).tolist()))
        pairs += sum(na[k] * nb[k] for k in na if k in nb)
    res = m.SearchByBoWBatch(0, A, B)
    t0 = time.perf_counter()
    for _ in range(reps):
        res = m.SearchByBoWBatch(0, A, B)
    dt = (time.perf_counter() - t0) / reps
    out["search_by_bow"] = {"calls_per_s": len(A) / dt, "hamming_pairs_per_s": pairs / dt, "ms_per_batch": dt * 1e3,
                            "batch": len(A), "matches_per_call": float(np.mean([r[1] for r in res]))}
    # ---- L1 score, one query against all candidates (KeyFrameDatabase.cc:238)
    cvecs = [(b[0], b[1]) for b in cb] * 8
    gvoc.score_batch((qb[0], qb[1]), cvecs)
    t0 = time.perf_counter()
    for _ in range(reps):
        sc = gvoc.score_batch((qb[0], qb[1]), cvecs)
    dt = (time.perf_counter() - t0) / reps
    out["l1_score"] = {"scores_per_s": len(cvecs) / dt, "ms_per_batch": dt * 1e3, "batch": len(cvecs)}
    if with_cpu:
        t0 = time.perf_counter(); oq = ovoc.transform(base, levelsup); dt = time.perf_counter() - t0
        out["transform"]["cpu_descriptors_per_s"] = n / dt
        oA = [M.Side(c, b[2], b[3], b[4], valid=valid, angles=ang) for c, b in zip(cands, cb)]
        oB = M.Side(base, qb[2], qb[3], qb[4], angles=ang)
        t0 = time.perf_counter()
        for a_ in oA:
            M.search_by_bow(0, a_, oB, 0.75, True)
        dt = time.perf_counter() - t0
        out["search_by_bow"]["cpu_calls_per_s"] = len(oA) / dt
        out["search_by_bow"]["cpu_hamming_pairs_per_s"] = pairs / dt
        t0 = time.perf_counter()
        for c in cvecs:
            M.bow_score_l1(qb[0], qb[1], c[0], c[1])
        dt = time.perf_counter() - t0
        out["l1_score"]["cpu_scores_per_s"] = len(cvecs) / dt
        out["cpu"] = "oracle port, 1 thread (includes the ctypes call overhead of one call per candidate)"
    # ---- SearchByProjection (TrackWithMotionModel / SearchLocalPoints: every tracked frame)
    from corb_slam_b200 import FrameView
    from corb_slam_b200.synth import projection_scene
    sc = projection_scene(1)
    c = sc["cur"]
    fv = FrameView(c["x"], c["y"], c["octave"], c["angle"], c["desc"], c["u_right"], sc["scales"], sc["bounds"], sc["K"], sc["mbf"],
                   sc["Tcw"], taken=sc["taken"])
    pm = ORBmatcher(0.8, True, device=device)
    last_args = (fv, sc["last_valid"], sc["Xw"], sc["mp_desc"], sc["last"]["octave"], sc["last"]["angle"], sc["Tlw"], 15.0)
    map_args = (fv, sc["in_view"], sc["proj"], sc["level"], sc["view_cos"], sc["mp_desc"])
    pm.SearchByProjectionLastFrame(*last_args, last_blocks=sc["last_blocks"]); pm.SearchByProjectionMapPoints(*map_args)
    t0 = time.perf_counter()
    for _ in range(reps):
        _, n_last = pm.SearchByProjectionLastFrame(*last_args, last_blocks=sc["last_blocks"])
    t1 = time.perf_counter()
    for _ in range(reps):
        _, n_map = pm.SearchByProjectionMapPoints(*map_args)
    t2 = time.perf_counter()
    out["search_by_projection"] = {"last_frame_calls_per_s": reps / (t1 - t0), "map_points_calls_per_s": reps / (t2 - t1),
                                   "last_frame_ms": 1e3 * (t1 - t0) / reps, "map_points_ms": 1e3 * (t2 - t1) / reps,
                                   "queries": int(len(sc["last_valid"])), "frame_features": int(fv.n), "matches": [int(n_last), int(n_map)],
                                   "note": "host arrays in, host match array out (H2D + two kernels + D2H per call)"}
    if with_cpu:
        from oracle import _proj_bind as PB
        cs = fv.c_struct()
        t0 = time.perf_counter()
        for _ in range(reps):
            PB.search_by_projection_last(cs, fv.n, sc["last_valid"], sc["last_blocks"], sc["Xw"], sc["mp_desc"], sc["last"]["octave"],
                                         sc["last"]["angle"], sc["Tlw"], 15.0, False, True)
        t1 = time.perf_counter()
        for _ in range(reps):
            PB.search_by_projection_map(cs, fv.n, sc["in_view"], None, sc["proj"], sc["level"], sc["view_cos"], sc["mp_desc"], 1.0, 0.8)
        t2 = time.perf_counter()
        out["search_by_projection"]["cpu_last_frame_calls_per_s"] = reps / (t1 - t0)
        out["search_by_projection"]["cpu_map_points_calls_per_s"] = reps / (t2 - t1)
    pm.close()
    out["pnp_ransac"] = bench_pnp(device, with_cpu)
    out["workload"] = "k=10 L=5 synthetic vocabulary, 2000 descriptors per frame, 32 candidate keyframes, level-2 node groups"
    m.close(); gvoc.close()
    return out


def pnp_workload(n_cand=20, n=150):
    """The relocalisation / map-fusion loop: n_cand candidate keyframes, each a PnPsolver over n matches (30 % of them wrong),
    SetRansacParameters(0.99, 10, 300, 4, 0.5, 5.991) as Tracking.cc:1414 / MapFusion.cpp:701, explicit draws."""
    import numpy as np
    from corb_slam_b200.synth import pnp_problem
    rng = np.random.default_rng(11)
    cands = []
    for c in range(n_cand):
        p = pnp_problem(500 + c, n=n, outlier_fraction=0.3 if c % 4 else 0.97, pixel_noise=0.5)  # every fourth candidate is a false positive
        draws = np.stack([rng.integers(0, n - k, 300) for k in range(4)], 1).astype(np.int32)
        cands.append((p, draws))
    return cands


def bench_pnp(device, with_cpu=True, reps=20):
    """EPnP-RANSAC batched over candidates: ms per batch through the C ABI (host arrays in and out)
    at a relocalisation-sized batch (20 candidates) and a map-fusion-sized one (200), beside the sequential oracle port on
    one host thread."""
    import numpy as np
    from corb_slam_b200 import PnPsolver

    def solvers(cands):
        out = []
        for p, draws in cands:
            n = len(p["p2d"])
            s = PnPsolver(p["p2d"], np.zeros(n, np.int64), [1.0], p["K"], p["p3d"], np.ones(n, bool), device=device)
            s.mvSigma2 = p["sigma2"]
            s.SetRansacParameters(0.99, 10, 300, 4, 0.5, 5.991)
            s.set_draws(draws)
            out.append(s)
        return out

    def gpu(cands):
        PnPsolver.iterate_batch(solvers(cands), 5)
        t = 0.0
        for _ in range(reps):
            ss = solvers(cands)
            packed = PnPsolver.pack_batch(ss, 5)
            t0 = time.perf_counter()
            PnPsolver.call_batch(device, packed)  # the C-ABI call: host arrays in, host results out
            t += time.perf_counter() - t0
        res = [s._unpack(packed[2][i], packed[4][i]) for i, s in enumerate(ss)]
        hyp = sum(s.mRansacMaxIts for s in ss)
        return {"candidates": len(cands), "hypotheses_per_batch": hyp, "ms_per_batch": 1e3 * t / reps, "hypotheses_per_s": hyp * reps / t,
                "poses_found": sum(r[0] is not None for r in res)}

    cands = pnp_workload()
    out = gpu(cands)
    out["matches_per_candidate"] = len(cands[0][0]["p2d"])
    out["note"] = "one corb_pnp_iterate_batch call: every RANSAC hypothesis of every candidate is evaluated (H2D + 6 kernels + D2H)"
    out["large_batch"] = gpu(pnp_workload(n_cand=200))
    if with_cpu:
        from oracle import _pnp_bind as PB
        t = 0.0
        its = 0
        for _ in range(3):
            for (p, draws), s in zip(cands, solvers(cands)):
                o = PB.PnpSolver(p["p2d"], p["p3d"], s.mvMaxError, *[float(v) for v in p["K"]], s.mRansacMinInliers, s.mRansacMaxIts)
                t0 = time.perf_counter()
                o.iterate(5, draws)
                t += time.perf_counter() - t0
                its += o.iterations
        out["cpu_ms_per_batch"] = 1e3 * t / 3
        out["cpu_iterations_per_batch"] = its / 3
        out["cpu"] = "oracle port, 1 thread, sequential with the reference's early exit (fewer hypotheses than the GPU evaluates)"
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.gpus
    total = 0.0
    t_all = 0.0
    per_step = max(4, min(32, args.sample_frames // max(1, args.steps)))
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_extract_fps(2, clients=n)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        fps, dt = cpu_extract_fps(per_step, clients=n)
        frames += n * per_step
        t_all += dt
    fps = frames / t_all
    line = {
        "impl": "reference", "metric": "stereo_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "config#2: synthetic 1242x375 stereo, 2000 ORB features/frame, 8 levels, FAST 20/7",
                   "clients": n, "frames_per_step": per_step},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 2 * n, "kind": "port",
                         "sample": "%d stereo frames per client, %d client(s), L/R on two threads each (Frame.cc:78-81); "
                                   "oracle port because the reference needs OpenCV/Eigen/ROS to build" % (per_step * args.steps, n),
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_ba:
        line["ba"] = dict(cpu_ba(BA_P, BA_L), metric="global_ba_ms_per_10k_landmarks", higher_is_better=False, impl="reference")
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from corb_slam_b200 import ORBextractor, extract_stereo, extract_stereo_device, frame_stereo
    from corb_slam_b200.synth import stereo_frame, frame_seed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    exl = ORBextractor(*ORB_PARAMS, device=local)
    exr = ORBextractor(*ORB_PARAMS, device=local)
    exl.copy_outputs = exr.copy_outputs = False
    frames = [stereo_frame(frame_seed(i + 100 * rank)) for i in range(N_POOL)]
    pinned = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) for l, r in frames]
    dev = [(a.cuda(), b.cuda()) for a, b in pinned]
    npin = [(a.numpy(), b.numpy()) for a, b in pinned]  # numpy views of the page-locked frames
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    sl = torch.cuda.ExternalStream(exl.stream())
    sr = torch.cuda.ExternalStream(exr.stream())

    def step_device(i, ev0=None, ev1=None):
        l, r = dev[i % N_POOL]
        if ev0 is not None:
            ev0.record(sl)
            sr.wait_event(ev0)
        extract_stereo_device(exl, exr, l.data_ptr(), r.data_ptr(), W, H, W)
        if ev1 is not None:
            evr = torch.cuda.Event()
            evr.record(sr)
            sl.wait_event(evr)
            ev1.record(sl)

    def step_host(i, pyr=False):
        l, r = pinned[i % N_POOL]
        return extract_stereo(exl, exr, npin[i % N_POOL][0], npin[i % N_POOL][1], want_pyramid=pyr)

    # ---- warm-up
    for i in range(max(args.warmup, 3)):
        step_device(i)
        exl.sync(); exr.sync()
        step_host(i)
    # ---- HBM-resident throughput: per-step CUDA events, L2 flushed between steps
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    evs = []
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_device(i, e0, e1)
        evs.append((e0, e1))
        exl.sync(); exr.sync()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- end to end through the C ABI with host buffers (wall clock; H2D + D2H inside)
    barrier()
    t0 = time.perf_counter()
    nk = 0
    for i in range(args.steps):
        kl, kr = step_host(i)
        nk += len(kl[0]) + len(kr[0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(i, pyr=True)
    e2e_pyr_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    n_stereo = 0
    for i in range(args.steps):  # the stereo Frame constructor: ExtractORB x2 + ComputeStereoMatches, pyramids stay in HBM
        fr = frame_stereo(exl, exr, npin[i % N_POOL][0], npin[i % N_POOL][1], 386.1448, 386.1448 / 718.856)
        n_stereo += int((fr[2] >= 0).sum())
    e2e_frame_s = time.perf_counter() - t0
    # ---- several clients sharing one GPU (capacity, not the headline): C threads, each with its own handle pair,
    #      each running the same host-buffer e2e call; ctypes releases the GIL inside the C ABI
    multi = None
    if args.clients_per_gpu > 1:
        C_ = args.clients_per_gpu
        pairs = [(ORBextractor(*ORB_PARAMS, device=local), ORBextractor(*ORB_PARAMS, device=local)) for _ in range(C_)]
        for a_, b_ in pairs:
            a_.copy_outputs = b_.copy_outputs = False
            extract_stereo(a_, b_, npin[0][0], npin[0][1])
        start = threading.Barrier(C_ + 1)

        def client(idx):
            a_, b_ = pairs[idx]
            start.wait()
            for i in range(args.steps):
                extract_stereo(a_, b_, npin[(i + idx) % N_POOL][0], npin[(i + idx) % N_POOL][1])

        ths = [threading.Thread(target=client, args=(c,)) for c in range(C_)]
        [t_.start() for t_ in ths]
        start.wait()
        t0 = time.perf_counter()
        [t_.join() for t_ in ths]
        dt = time.perf_counter() - t0
        multi = {"clients_per_gpu": C_, "value": C_ * args.steps / dt, "unit": "frames/s",
                 "note": "aggregate e2e frames/s of %d independent clients (threads) sharing this GPU; not the headline" % C_}
        for a_, b_ in pairs:
            a_.close(); b_.close()
    clocks = sampler.stop() if sampler else None

    # ---- global BA (second half of the BASELINE.json metric): landmark-sharded over the ranks, reduced camera system
    #      all-reduced over NCCL from inside corb_ba_solve when more than one GPU is attached
    ba_info, ba_ms, ba_rms, ba_edges, ba_allreduce = None, 0.0, None, 0, None
    if not args.no_ba:
        from corb_slam_b200 import Optimizer, torch_allreduce
        from corb_slam_b200.synth import ba_problem, ba_shard
        prob = ba_problem(BA_P, BA_L, seed=7)
        ba_edges = len(prob["edge_pose"])
        mine = ba_shard(prob, rank, world) if world > 1 else prob
        cb = torch_allreduce() if world > 1 else None
        Optimizer.BundleAdjustment(ba_shard(ba_problem(50, 2000, seed=1), rank, world) if world > 1 else ba_problem(50, 2000, seed=1),
                                   2, bRobust=False, device=local, allreduce=cb)  # warm-up (context, NCCL channels)
        # second warm-up at the measured size, one LM iteration: sizes the library's per-device memory arena, as in a
        # server that runs global BA repeatedly (the timed call below still rebuilds every structure from the host arrays)
        Optimizer.BundleAdjustment(mine, 1, bRobust=False, device=local, allreduce=cb)
        barrier()
        t0 = time.perf_counter()
        ba_out, ba_info = Optimizer.BundleAdjustment(mine, BA_ITERS, bRobust=False, device=local, allreduce=cb)
        torch.cuda.synchronize()
        ba_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        if rank == 0 and world == 1:
            import oracle
            from oracle import _ba_bind as B
            oracle.lib()
            ba_rms = B.chi2(ba_out)[1]  # checker only: RMS reprojection error of the GPU solution
        if world > 1:
            # the one exchange step of the sharded BA on its own: all-reduce(sum, fp64) of [S | bschur] (36 doubles per block of
            # the reduced camera system + 6 per pose), timed with CUDA events, against the NVLink roofline
            n_red = 36 * int(ba_info["reduced_blocks"]) + 6 * BA_P
            buf = torch.zeros(n_red, dtype=torch.float64, device="cuda")
            for _ in range(3):
                dist.all_reduce(buf)
            torch.cuda.synchronize(); barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(20):
                dist.all_reduce(buf)
            ev1.record(); torch.cuda.synchronize()
            ar_ms = ev0.elapsed_time(ev1) / 20
            tt = torch.tensor([ar_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ba_allreduce = {"message_bytes": 8 * n_red, "ms": float(tt.item())}

    t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_pyr_s * 1e3, ba_ms, e2e_frame_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_pyr_ms, ba_ms, e2e_frame_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        # per-kernel times of one image (eager replay with events) for the roofline of the dominant kernel
        exl.extract_device(dev[0][0].data_ptr(), W, H, W)
        exl.sync()
        prof = exl.profile(reps=20)
        agg = {}
        for name, ms in prof:
            base = name.split("[")[0]
            agg[base] = agg.get(base, 0.0) + ms
        top = max(prof, key=lambda kv: kv[1])  # the single longest launch (per-level launches run concurrently in the graph)
        pk, pk_kind = peaks()
        achieved = ALGO_BYTES_PER_IMAGE / (top[1] * 1e-3) / 1e9
        fps = world * args.steps / (dev_ms * 1e-3)
        kp_per_frame = nk / max(1, args.steps)
        h2d = 2 * W * H
        d2h = int(kp_per_frame * 60) + 16
        cpu_fps, cpu_dt = cpu_extract_fps(args.sample_frames, clients=1)
        cpu_fps_frame, _ = cpu_extract_fps(max(8, args.sample_frames // 4), clients=1, stereo=True)
        line = {
            "metric": "stereo_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "config#2: synthetic 1242x375 stereo, 2000 ORB features/frame, 8 levels, FAST 20/7",
                       "clients_per_gpu": 1, "frame_pool": N_POOL, "l2": "flushed (256 MiB write) between timed steps",
                       "timing": "CUDA events per step on the extractor streams, sum over steps, max over ranks"},
            "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "note": "corb_orb_extract_pair on two handles: page-locked host images in, host keypoints+descriptors out"},
            "e2e_with_pyramid": {"value": world * args.steps / (e2e_pyr_ms * 1e-3), "unit": "frames/s",
                                 "d2h_bytes_per_step": d2h + 2 * 1441432, "ms_per_step": e2e_pyr_ms / args.steps,
                                 "note": "also copies mvImagePyramid to the host (needed while ComputeStereoMatches is on the CPU)"},
            "e2e_stereo_frame": {"value": world * args.steps / (e2e_frame_ms * 1e-3), "unit": "frames/s",
                                 "ms_per_step": e2e_frame_ms / args.steps, "stereo_matches_per_frame": n_stereo / max(1, args.steps),
                                 "d2h_bytes_per_step": d2h + 8 * int(kp_per_frame / 2),
                                 "note": "corb_frame_stereo: ExtractORB left+right and Frame::ComputeStereoMatches on the GPU; "
                                         "the pyramids never leave HBM"},
            "multi_client": multi,
            "gpu_launches": exl.launches_per_extract() * args.steps,  # one set of launches per stereo pair (grid z = 2)
            "tma": exl.uses_tma(),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": ncu_traffic(top[0]), "peak_source": pk_kind,
                         "traffic_source": "profiles/r01b_stereo_frame_ncu_full.csv (ncu --set full, dram read + write of that launch / 2 images)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_IMAGE, "kernel_ms": top[1],
                         "whole_frame_frac": 2 * ALGO_BYTES_PER_IMAGE * fps / world / 1e9 / pk["hbm_gbs"],
                         "per_kernel_ms_sum_over_levels": {k: round(v, 5) for k, v in agg.items()},
                         "per_launch_ms": {k: round(v, 5) for k, v in prof}},
            "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": 2, "kind": "port",
                             "sample": "%d stereo frames, oracle port, L/R on two threads (Frame.cc:78-81), %.1f s"
                                       % (args.sample_frames, cpu_dt), "host_cores_available": os.cpu_count(),
                             "stereo_frame_value": cpu_fps_frame},
            "keypoints_per_frame": kp_per_frame,
        }
        if not args.no_matcher:
            line["matcher"] = bench_matcher(local)
        if ba_info is not None:
            ba_bytes = ba_algorithmic_bytes(ba_edges, BA_L, BA_P, ba_info["reduced_blocks"]) * ba_info["iterations"]
            ach = ba_bytes / (ba_ms * 1e-3) / 1e9
            line["ba"] = {
                "metric": "global_ba_ms_per_10k_landmarks", "value": ba_ms / (BA_L / 1e4), "unit": "ms/10k landmarks",
                "higher_is_better": False, "n_gpus": world, "scaling": "strong", "dtype": "f64",
                "config": {"workload": "synthetic street BA: P=%d keyframes, L=%d landmarks, E=%d observations (80%% stereo), "
                                       "%d LM iterations, bRobust=false, seed 7" % (BA_P, BA_L, ba_edges, BA_ITERS),
                           "sharding": "landmarks l %% N per rank, poses replicated, all-reduce(sum, fp64) of the reduced "
                                       "camera system per LM trial" if world > 1 else "single GPU, no collective"},
                "ms_total": ba_ms, "ms_in_library": ba_info["ms_total"], "ms_reduced_camera_solves": ba_info["ms_solve"],
                "iterations": ba_info["iterations"], "trials": ba_info["n_trials"], "trial_accepted": ba_info["trial_accepted"],
                "chi2_initial": ba_info["chi2_initial"], "chi2_final": ba_info["chi2_final"], "rms_px": ba_rms,
                "reduced_blocks": ba_info["reduced_blocks"],
                "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                             "algorithmic_bytes_per_iteration": ba_bytes // max(1, ba_info["iterations"]),
                             "note": "whole-call wall time incl. host structure build and H2D/D2H; the reduced-camera solve "
                                     "(chunked block-skyline Cholesky) is bound by the serial latency of a block column"},
                "cpu_baseline": cpu_ba(BA_P, BA_L) if not args.no_cpu_ba else None,
            }
            if ba_allreduce is not None:
                S, ms = ba_allreduce["message_bytes"], ba_allreduce["ms"]
                bus = 2.0 * (world - 1) / world * S / (ms * 1e-3) / 1e9  # bytes every GPU sends + receives per second
                line["ba"]["allreduce"] = {
                    "bound": "nvlink", "message_bytes": S, "ms": ms, "per_call_ms_total": ms * ba_info["n_trials"],
                    "achieved": bus, "peak": 900.0, "unit": "GB/s", "frac": bus / 900.0,
                    "note": "NCCL all-reduce(sum, fp64) of the reduced camera system [S | bschur], one per LM trial; bus bandwidth "
                            "2(n-1)/n * S / t against 900 GB/s per direction (NVLink 5); a %.1f MB message is latency bound" % (S / 1e6)}
        print(json.dumps(line))
    exl.close(); exr.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sample-frames", type=int, default=200, help="stereo frames of the bounded CPU baseline sample")
    ap.add_argument("--no-ba", action="store_true", help="skip the global-BA half of the metric")
    ap.add_argument("--no-cpu-ba", action="store_true", help="skip the CPU BA baseline (about 10 s)")
    ap.add_argument("--clients-per-gpu", type=int, default=4, help="extra capacity measurement with this many clients on one GPU (1 = skip)")
    ap.add_argument("--no-matcher", action="store_true", help="skip the SearchByBoW / vocabulary / score section")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
