#!/usr/bin/env python
"""Benchmark of the CORB-SLAM hot path on B200: synthetic 1242x375 stereo frames/s (BASELINE.json metric, config #2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun for N > 1); every rank is one `corbslam_client` stream pinned to its GPU (replicas:
frames of different robots are independent), so scaling is weak and there is no data-path collective.
A STEP = FRAMES_PER_STEP (128) stereo frames, in both arms; a frame = ORBextractor::operator() on the left and the right
image (Frame.cc:78-81). The client keeps IN_FLIGHT = 8 frames in flight (that many handle pairs, corb_orb_extract_pair_submit/_wait):
frame i + 1 is extracted while the tracking thread would consume frame i.

  value : frames/s with the images already resident in HBM (a pool of 160 distinct stereo pairs = 149 MB, larger than the
          126 MB L2, walked cyclically; L2 also flushed between steps) and results left in HBM; CUDA events around every
          step on the extractor streams, max over ranks
  e2e   : frames/s through the reference-facing C-ABI calls with HOST buffers (page-locked images read over PCIe, keypoints +
          descriptors back to the host, inside the timed region; wall clock)
  roofline : the whole stereo frame against the measured HBM copy bandwidth (algorithmic bytes of SURVEY.md section 8d), plus
          the per-kernel table with each kernel's own algorithmic bytes and its duration measured live (CUDA events) and
          in the committed ncu capture (profiles/r02_*)
  --impl reference : the reference's own ORBextractor.cc (oracle/_ref/libref.so, compiled unmodified from the reference tree;
          the oracle port if that library was not shipped), left and right image on two threads per client like Frame.cc:78-81.
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
N_BASE = 16           # distinct synthetic scenes ...
N_POOL = 160          # ... shifted into 160 distinct stereo pairs: 160 x 2 x 465 750 B = 149 MB > 126 MB of L2
FRAMES_PER_STEP = 128  # one step = 128 stereo frames, in both arms (20 steps = 2 560 frames = ~75 ms of GPU time)
IN_FLIGHT = int(os.environ.get("CORB_BENCH_IN_FLIGHT", "8"))  # stereo frames a client keeps in flight (handle pairs)
ALGO_BYTES_PER_IMAGE = W * H + 1441432 + 60 * 2000  # input + pyramid + 60 B per keypoint (K = 2000)
WORKLOAD = "config#2: synthetic 1242x375 stereo, 2000 ORB features/frame, 8 levels, FAST 20/7"


def bench_config():
    """The same dictionary in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "frames_per_step": FRAMES_PER_STEP, "clients_per_gpu": 1, "frames_in_flight": IN_FLIGHT,
            "frame_pool": "%d distinct stereo pairs (%d MB, larger than L2), walked cyclically" % (N_POOL, N_POOL * 2 * W * H // 1000000),
            "l2": "inputs larger than L2 and a 256 MiB flush between timed steps",
            "timing": "CUDA events around every step on the extractor streams, sum over steps, max over ranks"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def frame_pool(rank):
    """N_POOL distinct stereo pairs: N_BASE generated scenes, each also rolled by multiples of 37 columns."""
    from corb_slam_b200.synth import stereo_frame, frame_seed
    base = [stereo_frame(frame_seed(i + 100 * rank)) for i in range(N_BASE)]
    out = []
    for k in range(N_POOL):
        l, r = base[k % N_BASE]
        sh = 37 * (k // N_BASE)
        out.append((np.ascontiguousarray(np.roll(l, sh, axis=1)), np.ascontiguousarray(np.roll(r, sh, axis=1))) if sh else (l, r))
    return out


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


def _ncu_rows(path):
    import csv
    if not os.path.exists(path):
        return None, None
    rows = [r for r in csv.reader(open(path)) if len(r) > 4]
    return rows[0], rows[1:]


def ncu_frame_capture():
    """Per-kernel duration and DRAM bytes of ONE stereo frame from the committed `ncu --set full` capture
    (profiles/r02_stereo_frame_ncu_full.csv, written by tools/summarize_ncu.py; every launch covers both images of the pair).
    -> {kernel label: {"us": duration, "dram_bytes": read + write}} or {} when the capture is missing."""
    hdr, rows = _ncu_rows(os.path.join(ROOT, "profiles", "r02_stereo_frame_ncu_full.csv"))
    if hdr is None:
        hdr, rows = _ncu_rows(os.path.join(ROOT, "profiles", "r01c_stereo_frame_ncu_full.csv"))
    if hdr is None:
        return {}, None

    def col(prefix):
        hits = [i for i, h in enumerate(hdr) if h.startswith(prefix)]
        return hits[0] if hits else None

    def scaled(v, h):
        unit = h[h.index("[") + 1:h.index("]")] if "[" in h else ""
        return float(v) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3,
                           "msecond": 1e3}.get(unit, 1.0)
    cd, cr, cw = col("gpu__time_duration.sum"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    out, seen = {}, {}
    for r in rows:
        base = r[0].split("(")[0].replace("corb::", "").strip()
        k = seen.get(base, 0)
        seen[base] = k + 1
        if base == "k_resize":
            k += 1
        label = "%s[%d]" % (base, k) if base in ("k_resize", "k_fast_cells", "k_octtree") else base
        if label in out:
            continue  # the first frame of the capture only
        try:
            out[label] = {"us": scaled(r[cd], hdr[cd]) if cd is not None else None,
                          "dram_bytes": (scaled(r[cr], hdr[cr]) + scaled(r[cw], hdr[cw])) if cr is not None and cw is not None else None}
        except (ValueError, IndexError):
            pass
    return out, "profiles/r02_stereo_frame_ncu_full.csv" if os.path.exists(os.path.join(ROOT, "profiles", "r02_stereo_frame_ncu_full.csv")) \
        else "profiles/r01c_stereo_frame_ncu_full.csv"


def kernel_algorithmic_bytes(ex, counts):
    """Each kernel's OWN algorithmic bytes for one image. counts: per-level (candidates, kept)."""
    sizes = [ex.level_size(l, W, H) for l in range(ex.nlevels)]
    px = [w * h for w, h in sizes]
    out = {"k_import": 2 * px[0], "k_blur": 2 * sum(px), "k_orient_desc": sum(k for _, k in counts) * (961 + 512 + 60)}
    for l in range(ex.nlevels):
        if l:
            out["k_resize[%d]" % l] = px[l - 1] + px[l]
        out["k_fast_cells[%d]" % l] = px[l] + 8 * counts[l][0]
        out["k_octtree[%d]" % l] = 8 * (counts[l][0] + counts[l][1])
    return out


def reference_extractors(n):
    """`n` (left, right) extractor pairs of the CPU arm: the reference's own ORBextractor.cc when oracle/_ref/libref.so is
    there ("reference"), else the oracle port ("port")."""
    from oracle import ref
    if ref.available():
        return [(ref.ORBextractor(*ORB_PARAMS), ref.ORBextractor(*ORB_PARAMS)) for _ in range(n)], "reference"
    import oracle
    return [(oracle.OrbExtractor(*ORB_PARAMS), oracle.OrbExtractor(*ORB_PARAMS)) for _ in range(n)], "port"


def cpu_extract_fps(n_frames, clients=1, stereo=False, exs=None):
    """The reference path on the host cores: `clients` concurrent clients, each running left / right on two threads
    (Frame.cc:78-81); with stereo=True also Frame::ComputeStereoMatches (Frame.cc:90) on the calling thread (oracle port).
    -> (frames/s, seconds, kind)."""
    from corb_slam_b200.synth import stereo_frame, frame_seed
    frames = [stereo_frame(frame_seed(i)) for i in range(min(n_frames, N_BASE))]
    kind = "port"
    if stereo:
        import oracle
        exs = [(oracle.OrbExtractor(*ORB_PARAMS), oracle.OrbExtractor(*ORB_PARAMS)) for _ in range(clients)]
    elif exs is None:
        exs, kind = reference_extractors(clients)
    else:
        exs, kind = exs
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
    return clients * n_frames / dt, dt, kind


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


def vocabulary_text(stripped=False):
    """The reference's ORBvoc.txt (corbslam_client/Vocabulary/ORBvoc.txt.tar.gz; a copy travels in oracle/_ref/), untarred once
    into a temp directory. stripped: without the trailing newline (what the reference's own loader needs, see oracle/ref.py)."""
    import tarfile
    import tempfile
    cache = os.path.join(tempfile.gettempdir(), "corb_voc_%d" % os.getuid())
    full = os.path.join(cache, "ORBvoc.txt")
    if not os.path.exists(full):
        src = [q for q in (os.path.join(ROOT, "oracle", "_ref", "ORBvoc.txt.tar.gz"),
                           "/root/reference/corbslam_client/Vocabulary/ORBvoc.txt.tar.gz") if os.path.exists(q)]
        if not src:
            return None
        os.makedirs(cache, exist_ok=True)
        with tarfile.open(src[0]) as t:
            t.extract("ORBvoc.txt", cache + ".tmp", filter="data")
        os.replace(os.path.join(cache + ".tmp", "ORBvoc.txt"), full)
    if not stripped:
        return full
    cut = os.path.join(cache, "ORBvoc_stripped.txt")
    if not os.path.exists(cut):
        data = open(full, "rb").read().rstrip()
        with open(cut + ".tmp", "wb") as f:
            f.write(data)
        os.replace(cut + ".tmp", cut)
    return cut


def matcher_workload(device, n_cand=32):
    """Matching / BoW workload on the REAL vocabulary: the query is a bench frame, the `n_cand` candidate keyframes are the
    same scene seen again (shifted by a few pixels, sensor noise) - descriptors extracted by the GPU extractor."""
    from corb_slam_b200 import ORBextractor
    from corb_slam_b200.synth import stereo_frame, frame_seed
    ex = ORBextractor(*ORB_PARAMS, device=device)
    base = stereo_frame(frame_seed(0))[0]
    rng = np.random.default_rng(3)
    q = tuple(a.copy() for a in ex(base))
    cands = []
    for c in range(n_cand):
        img = np.roll(base, int(rng.integers(-12, 13)), axis=1).astype(np.int16) + rng.normal(0, 2.5, base.shape).round().astype(np.int16)
        cands.append(tuple(a.copy() for a in ex(img.clip(0, 255).astype(np.uint8))))
    ex.close()
    return q, cands


def bench_matcher(device, with_cpu=True):
    """Calls/s and Hamming pairs/s of SearchByBoW (batched over candidates), descriptors/s of the vocabulary descent,
    scores/s of the L1 BoW score on ORBvoc.txt (k = 10, L = 6, 1 082 073 nodes, levelsup = 4 as Frame.cc:404); each beside the
    reference's own DBoW2 / ORBmatcher.cc (oracle/_ref) on one host thread."""
    from corb_slam_b200 import BowFeatures, ORBmatcher, ORBVocabulary
    path = vocabulary_text()
    if path is None:
        return {"unavailable": "ORBvoc.txt.tar.gz was not shipped (oracle/_ref/)"}
    (qk, base), cand_kd = matcher_workload(device)
    cands = [d for _, d in cand_kd]
    gvoc = ORBVocabulary(device=device)
    t0 = time.perf_counter()
    if not gvoc.loadFromTextFile(path):
        return {"unavailable": "corb_voc_load_text failed"}
    load_s = time.perf_counter() - t0
    levelsup = 4
    n = len(base)
    rng = np.random.default_rng(0)
    ang = qk["angle"]
    out = {"vocabulary": {"k": gvoc.k, "L": gvoc.L, "nodes": gvoc.n_nodes, "words": gvoc.n_words, "load_s": load_s}}
    # ---- transform (voc->transform, Frame.cc:404)
    gvoc.transform(base, levelsup)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        qb = gvoc.transform(base, levelsup)
    dt = (time.perf_counter() - t0) / reps
    out["transform"] = {"descriptors_per_s": n / dt, "ms_per_call": dt * 1e3, "descriptors": n,
                        "roofline": {"bound": "hbm", "algorithmic_bytes": 48 * n, "achieved": 48 * n / dt / 1e9, "unit": "GB/s",
                                     "note": "32 B in + 16 B out per descriptor (SURVEY.md section 8d); the call is host-overhead bound"}}
    cb = [gvoc.transform(c, levelsup) for c in cands]
    # ---- SearchByBoW, batched over candidates (MapFusion.cpp:691 / Tracking.cc:1405)
    m = ORBmatcher(0.75, True, device=device)
    valids = [(rng.random(len(c)) < 0.7).astype(np.uint8) for c in cands]
    A = [BowFeatures(c, b[2], b[3], b[4], valid=v, angles=k["angle"]) for (k, c), b, v in zip(cand_kd, cb, valids)]
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
    bow_bytes = sum(36 * (len(c) + n) + 4 * n for c in cands)
    out["search_by_bow"] = {"calls_per_s": len(A) / dt, "hamming_pairs_per_s": pairs / dt, "ms_per_batch": dt * 1e3,
                            "batch": len(A), "matches_per_call": float(np.mean([r[1] for r in res])),
                            "roofline": {"bound": "hbm", "algorithmic_bytes": bow_bytes, "achieved": bow_bytes / dt / 1e9, "unit": "GB/s",
                                         "note": "36 (N_A + N_B) + 4 N_out bytes per call (SURVEY.md section 8d)"}}
    # ---- L1 score, one query against all candidates (KeyFrameDatabase.cc:238)
    cvecs = [(b[0], b[1]) for b in cb] * 8
    gvoc.score_batch((qb[0], qb[1]), cvecs)
    t0 = time.perf_counter()
    for _ in range(reps):
        sc = gvoc.score_batch((qb[0], qb[1]), cvecs)
    dt = (time.perf_counter() - t0) / reps
    out["l1_score"] = {"scores_per_s": len(cvecs) / dt, "ms_per_batch": dt * 1e3, "batch": len(cvecs)}
    # ---- the same three operations on device-resident records (corb_frame_bow / corb_bow_match_stores / corb_bow_score_stores):
    #      descriptors and BoW / feature vectors never leave HBM, only liveness masks go up and match arrays / scores come down
    from corb_slam_b200 import BowRecord, ORBextractor
    from corb_slam_b200.synth import stereo_frame, frame_seed
    ex = ORBextractor(*ORB_PARAMS, device=device)
    img0 = stereo_frame(frame_seed(0))[0]
    ex(img0)
    qrec = BowRecord(2048, device=device).from_extractor(ex, gvoc, levelsup)
    t0 = time.perf_counter()
    for _ in range(reps):
        qrec.from_extractor(ex, gvoc, levelsup)  # enqueues; the records are built back to back on the extractor's stream
    qrec.sync()
    dt = (time.perf_counter() - t0) / reps
    got = qrec.download()
    assert all(a.tobytes() == b.tobytes() for a, b in zip(got, qb)), "device-built BoW differs from corb_voc_transform"
    crecs = []
    import torch
    for (ck, cd) in cand_kd:
        dd = torch.from_numpy(np.ascontiguousarray(cd)).cuda()
        dk = torch.from_numpy(np.ascontiguousarray(ck).view(np.uint8)).cuda()
        crecs.append(BowRecord(2048, device=device).from_device(gvoc, dd.data_ptr(), len(cd), levelsup, d_kps=dk.data_ptr()))
    rres = m.SearchByBoWRecords(0, crecs, valids, [qrec] * len(crecs))
    assert all(g[1] == r_[1] and np.array_equal(g[0], r_[0]) for g, r_ in zip(res, rres)), "record SearchByBoW differs from the host-array path"
    t0 = time.perf_counter()
    for _ in range(reps):
        m.SearchByBoWRecords(0, crecs, valids, [qrec] * len(crecs))
    dt_m = (time.perf_counter() - t0) / reps
    rsc = BowRecord.score(gvoc, qrec, crecs * 8)
    assert rsc.tobytes() == np.asarray(sc).tobytes(), "record L1 scores differ from the host-array path"
    t0 = time.perf_counter()
    for _ in range(reps):
        BowRecord.score(gvoc, qrec, crecs * 8)
    dt_s = (time.perf_counter() - t0) / reps
    out["device_records"] = {
        "frame_bow_us": dt * 1e6, "frame_bow_descriptors_per_s": n / dt,
        "search_by_bow": {"ms_per_batch": dt_m * 1e3, "calls_per_s": len(crecs) / dt_m, "hamming_pairs_per_s": pairs / dt_m,
                          "roofline": {"bound": "hbm", "algorithmic_bytes": bow_bytes, "achieved": bow_bytes / dt_m / 1e9, "unit": "GB/s"}},
        "l1_score": {"ms_per_batch": dt_s * 1e3, "scores_per_s": len(crecs) * 8 / dt_s},
        "note": "corb_frame_bow (transform + BowVector / FeatureVector build on the device, one 16-byte D2H), "
                "corb_bow_match_stores, corb_bow_score_stores; results asserted equal to the host-array entry points"}
    for r_ in crecs + [qrec]:
        r_.close()
    ex.close()
    if with_cpu:
        from oracle import ref
        from oracle import _match_bind as M
        if ref.available():
            rvoc = ref.ORBVocabulary(vocabulary_text(stripped=True))
            cpu_transform, cpu_score, cpu_bow, kind = rvoc.transform, rvoc.score, ref.search_by_bow, "reference"
        else:
            import oracle
            oracle.lib()
            ovoc = M.Vocabulary.load_text(path)
            cpu_transform, cpu_score, cpu_bow, kind = ovoc.transform, M.Vocabulary.score, M.search_by_bow, "port"
        t0 = time.perf_counter(); oq = cpu_transform(base, levelsup); dt = time.perf_counter() - t0
        assert all(x.tobytes() == y.tobytes() for x, y in zip(oq, qb)), "transform differs from the CPU reference"
        out["transform"]["cpu_descriptors_per_s"] = n / dt
        oA = [M.Side(c, b[2], b[3], b[4], valid=v, angles=k["angle"]) for (k, c), b, v in zip(cand_kd, cb, valids)]
        oB = M.Side(base, qb[2], qb[3], qb[4], angles=ang)
        t0 = time.perf_counter()
        cres = [cpu_bow(0, a_, oB, 0.75, True) for a_ in oA]
        dt = time.perf_counter() - t0
        assert all(g[1] == c[1] and np.array_equal(g[0], c[0]) for g, c in zip(res, cres)), "SearchByBoW differs from the CPU reference"
        out["search_by_bow"]["cpu_calls_per_s"] = len(oA) / dt
        out["search_by_bow"]["cpu_hamming_pairs_per_s"] = pairs / dt
        t0 = time.perf_counter()
        csc = [cpu_score((qb[0], qb[1]), c) for c in cvecs]
        dt = time.perf_counter() - t0
        assert np.asarray(csc).tobytes() == np.asarray(sc).tobytes(), "L1 scores differ from the CPU reference"
        out["l1_score"]["cpu_scores_per_s"] = len(cvecs) / dt
        out["cpu"] = {"kind": kind, "threads": 1,
                      "note": "the reference's own DBoW2 / ORBmatcher.cc (oracle/_ref/libref.so)" if kind == "reference" else "oracle port"}
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
    out["workload"] = "ORBvoc.txt (k=10, L=6, 1 082 073 nodes, levelsup 4), %d descriptors of a bench frame, 32 candidate keyframes = the scene seen again" % n
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
    cores = os.cpu_count() or 2
    # concurrent reference clients on ALL the host cores, two threads each (Frame.cc:78-81): the reference's own code with all
    # the host threads it can use (our arm keeps IN_FLIGHT frames per GPU in flight; one client alone is reported beside it)
    clients = max(1, cores // 2)
    per_client = max(1, (n * FRAMES_PER_STEP) // clients)
    exs = reference_extractors(clients)
    cpu_extract_fps(2, clients=clients, exs=exs)  # warm-up
    t_all, frames, per = 0.0, 0, []
    for _ in range(args.steps):
        fps, dt, kind = cpu_extract_fps(per_client, clients=clients, exs=exs)
        frames += clients * per_client
        t_all += dt
        per.append(dt)
    one_fps = cpu_extract_fps(16, clients=1, exs=(exs[0][:1], exs[1]))[0]  # one client alone, as the reference runs it
    fps = frames / t_all
    line = {
        "impl": "reference", "metric": "stereo_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": bench_config(),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 2 * clients, "kind": kind, "cpu_model": cpu_model(),
                         "one_client_value": one_fps,
                         "sample": "%d steps x %d stereo frames per client, %d concurrent client(s), L/R on two threads each (Frame.cc:78-81); %s"
                                   % (args.steps, per_client, clients,
                                      "the reference's own ORBextractor.cc compiled unmodified (oracle/_ref/libref.so)" if kind == "reference"
                                      else "oracle port (oracle/_ref/libref.so was not shipped)"),
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_ba:
        line["ba"] = dict(cpu_ba(BA_P, BA_L), metric="global_ba_ms_per_10k_landmarks", higher_is_better=False, impl="reference")
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from corb_slam_b200 import (ORBextractor, extract_stereo, extract_stereo_device, extract_stereo_submit, extract_stereo_wait,
                                frame_stereo_submit, frame_stereo_wait)

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

    # IN_FLIGHT handle pairs = that many stereo frames in flight
    NP = IN_FLIGHT
    pairs = [(ORBextractor(*ORB_PARAMS, device=local), ORBextractor(*ORB_PARAMS, device=local)) for _ in range(NP)]
    for a_, b_ in pairs:
        a_.copy_outputs = b_.copy_outputs = False
        a_.set_host_transfer(1); b_.set_host_transfer(1)  # frames in flight: host images through the copy engines
    # the blocking call (one frame in flight) keeps the default transfer: the import kernel reads the images in place
    exl, exr = ORBextractor(*ORB_PARAMS, device=local), ORBextractor(*ORB_PARAMS, device=local)
    exl.copy_outputs = exr.copy_outputs = False
    frames = frame_pool(rank)
    pinned = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) for l, r in frames]
    dev = [(a.cuda(), b.cuda()) for a, b in pinned]
    npin = [(a.numpy(), b.numpy()) for a, b in pinned]  # numpy views of the page-locked frames
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    streams = [torch.cuda.ExternalStream(p_[0].stream()) for p_ in pairs]  # a pair runs on its left handle's stream
    F = FRAMES_PER_STEP
    mbf, mb = 386.1448, 386.1448 / 718.856

    def step_device(step, ev0, ev1):
        """F stereo frames, round-robin over the IN_FLIGHT handle pairs; events bracket the whole step on all their streams."""
        ev0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(ev0)
        for f in range(F):
            i = (step * F + f) % N_POOL
            a_, b_ = pairs[f % NP]
            extract_stereo_device(a_, b_, dev[i][0].data_ptr(), dev[i][1].data_ptr(), W, H, W)
        for st in streams[1:]:
            evb = torch.cuda.Event()
            evb.record(st)
            streams[0].wait_event(evb)
        ev1.record(streams[0])

    def step_host(step, submit, wait):
        """F stereo frames through the split C-ABI calls from this one thread, IN_FLIGHT in flight; returns keypoints seen."""
        nk = 0
        base = step * F
        for f in range(min(NP - 1, F)):
            submit(*pairs[f % NP], *npin[(base + f) % N_POOL])
        for f in range(F):
            if f + NP - 1 < F:
                submit(*pairs[(f + NP - 1) % NP], *npin[(base + f + NP - 1) % N_POOL])
            res = wait(*pairs[f % NP])
            nk += len(res[0][0]) + len(res[1][0])
        return nk

    sub_frame = lambda a_, b_, l, r: frame_stereo_submit(a_, b_, l, r, mbf, mb)
    # ---- warm-up (plans, graphs, page-locked buffers of both pairs)
    for w_ in range(max(args.warmup, 3)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_device(w_, e0, e1)
        torch.cuda.synchronize()
    step_host(0, extract_stereo_submit, extract_stereo_wait)
    step_host(0, sub_frame, frame_stereo_wait)
    extract_stereo(exl, exr, *npin[0])
    # ---- HBM-resident throughput
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    evs = []
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_device(i, e0, e1)
        evs.append((e0, e1))
        torch.cuda.synchronize()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = sum(step_ms)
    # ---- latency of ONE frame alone on the GPU (no second frame in flight), CUDA events per frame
    lat = []
    lat_stream = torch.cuda.ExternalStream(exl.stream())  # the pair runs on its left handle's stream
    for i in range(min(200, 4 * F)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lat_stream)
        extract_stereo_device(exl, exr, dev[i % N_POOL][0].data_ptr(), dev[i % N_POOL][1].data_ptr(), W, H, W)
        e1.record(lat_stream)
        torch.cuda.synchronize()
        lat.append(1e3 * e0.elapsed_time(e1))
    # ---- end to end through the C ABI with host buffers (wall clock; H2D + D2H inside), IN_FLIGHT frames in flight
    barrier()
    t0 = time.perf_counter()
    nk = 0
    for i in range(args.steps):
        nk += step_host(i, extract_stereo_submit, extract_stereo_wait)
    e2e_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):  # the same frames through the blocking call (one frame in flight)
        for f in range(F):
            extract_stereo(exl, exr, *npin[(i * F + f) % N_POOL])
    e2e_block_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    n_stereo = 0
    for i in range(args.steps):  # the stereo Frame constructor: ExtractORB x2 + ComputeStereoMatches, pyramids stay in HBM
        base = i * F
        for f in range(min(NP - 1, F)):
            sub_frame(*pairs[f % NP], *npin[(base + f) % N_POOL])
        for f in range(F):
            if f + NP - 1 < F:
                sub_frame(*pairs[(f + NP - 1) % NP], *npin[(base + f + NP - 1) % N_POOL])
            fr = frame_stereo_wait(*pairs[f % NP])
            n_stereo += int((fr[2] >= 0).sum())
    e2e_frame_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    n_frames = args.steps * F

    # ---- global BA (second half of the BASELINE.json metric): landmark-sharded over the ranks, reduced camera system
    #      all-reduced over NCCL from inside corb_ba_solve when more than one GPU is attached
    ba_info, ba_ms, ba_rms, ba_edges, ba_allreduce, ba_parity, ba_server = None, 0.0, None, 0, None, None, None
    if not args.no_ba:
        from corb_slam_b200 import Optimizer, page_locked, torch_allreduce
        from corb_slam_b200.synth import ba_problem, ba_shard
        prob = ba_problem(BA_P, BA_L, seed=7)
        ba_edges = len(prob["edge_pose"])
        # the flat edge arrays live in page-locked memory, like the buffers the C++ shim flattens the graph into (shim/Optimizer_gba.h)
        mine = page_locked(ba_shard(prob, rank, world) if world > 1 else prob)
        cb = torch_allreduce(device=local) if world > 1 else None
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
            # multi-GPU parity inside the driver-run bench: rank spread of the replicated poses, and on rank 0 the sharded
            # result against the unsharded solve on one GPU (accept / reject sequence, poses, landmarks)
            poses = torch.from_numpy(np.concatenate([ba_out["pose_q"], ba_out["pose_t"]], 1)).cuda()
            pmax, pmin = poses.clone(), poses.clone()
            dist.all_reduce(pmax, op=dist.ReduceOp.MAX); dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
            pts = torch.zeros((BA_L, 3), dtype=torch.float64, device="cuda")
            pts[torch.from_numpy(mine["_point_ids"]).cuda()] = torch.from_numpy(ba_out["point_xyz"]).cuda()
            dist.all_reduce(pts)
            if rank == 0:
                prob_pl = page_locked(prob)
                Optimizer.BundleAdjustment(prob_pl, 1, bRobust=False, device=local)
                t0 = time.perf_counter()
                full, finfo = Optimizer.BundleAdjustment(prob_pl, BA_ITERS, bRobust=False, device=local)
                single_ms = (time.perf_counter() - t0) * 1e3
                ba_parity = {"accept_sequence_equal": ba_info["trial_accepted"] == finfo["trial_accepted"],
                             "max_abs_pose_t": float(np.abs(ba_out["pose_t"] - full["pose_t"]).max()),
                             "max_abs_pose_q": float(np.abs(ba_out["pose_q"] - full["pose_q"]).max()),
                             "max_abs_point": float(np.abs(pts.cpu().numpy() - full["point_xyz"]).max()),
                             "rank_spread": float((pmax - pmin).abs().max()),
                             "chi2_final_single": finfo["chi2_final"], "ms_single_gpu_same_run": single_ms,
                             "speedup_vs_single_gpu_same_run": single_ms / ba_ms,
                             "strong_scaling_efficiency": single_ms / ba_ms / world}
            barrier()
            # the server's own shape of the same thing: ONE process, N threads, N ncclComm_t, ncclAllReduce hook in C
            # (tests/host_harness/ba_nccl.cpp = what corbslam_server's GBA thread would call); the other ranks wait on the
            # rendezvous store (a host-side wait: a NCCL barrier would spin on the GPUs the harness is using)
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                import tempfile
                from corb_slam_b200.ba_file import read_result, write_problem
                exe = os.path.join(ROOT, "tests", "host_harness", "ba_nccl")
                if not os.path.exists(exe):
                    subprocess.run(["make", "-C", os.path.dirname(exe)], capture_output=True)
                tmp = tempfile.mkdtemp()
                try:
                    write_problem(os.path.join(tmp, "p.bin"), prob)
                    r = subprocess.run([exe, str(world), os.path.join(tmp, "p.bin"), os.path.join(tmp, "r.bin"), str(BA_ITERS)],
                                       capture_output=True, text=True, timeout=600)
                    if r.returncode == 0:
                        res = read_result(os.path.join(tmp, "r.bin"), BA_P, BA_L)
                        ba_server = {"ms_total": res["ms_total"], "value": res["ms_total"] / (BA_L / 1e4), "unit": "ms/10k landmarks",
                                     "accept_sequence_equal_to_torchrun_path": res["trial_accepted"] == ba_info["trial_accepted"],
                                     "max_abs_pose_t_vs_torchrun_path": float(np.abs(res["pose_t"] - ba_out["pose_t"]).max()),
                                     "rank_spread": res["rank_spread"], "chi2_final": res["chi2_final"],
                                     "how": "one process, %d threads, %d ncclComm_t (ncclCommInitAll), ncclAllReduce hook" % (world, world)}
                    else:
                        ba_server = {"failed": (r.stdout + r.stderr)[-400:]}
                except Exception as ex:  # the harness is an extra measurement, never the headline
                    ba_server = {"failed": repr(ex)}
                store.set("corb_ba_server_done", "1")
            else:
                store.wait(["corb_ba_server_done"])
            barrier()
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

    t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_block_s * 1e3, ba_ms, e2e_frame_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_block_ms, ba_ms, e2e_frame_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        # per-kernel times of one image (eager replay with events between the launches) and the per-level counts
        exl.extract_device(dev[0][0].data_ptr(), W, H, W)
        exl.sync()
        counts = [(len(exl.tap_candidates(l)), exl.tap_level_count(l)) for l in range(exl.nlevels)]
        prof = dict(exl.profile(reps=20))
        kbytes = kernel_algorithmic_bytes(exl, counts)
        cap, cap_src = ncu_frame_capture()
        pk, pk_kind = peaks()
        fps = world * n_frames / (dev_ms * 1e-3)
        kernels = {}
        for name, nbytes in kbytes.items():
            ent = {"algorithmic_bytes": nbytes, "ms_live": prof.get(name)}
            if name in cap and cap[name]["us"]:
                n_img = 2  # every launch of the capture covers both images of the pair
                ent["us_ncu"] = cap[name]["us"]
                ent["frac_ncu"] = n_img * nbytes / (cap[name]["us"] * 1e-6) / 1e9 / pk["hbm_gbs"]
                ent["dram_bytes_ncu"] = cap[name]["dram_bytes"]
            kernels[name] = ent
        with_ncu = {k: v for k, v in kernels.items() if "us_ncu" in v}
        dom = max(with_ncu, key=lambda k: with_ncu[k]["us_ncu"]) if with_ncu else max(prof, key=prof.get)
        traffic = sum(v["dram_bytes_ncu"] for v in with_ncu.values() if v.get("dram_bytes_ncu")) if with_ncu else None
        achieved = 2 * ALGO_BYTES_PER_IMAGE * (fps / world) / 1e9
        kp_per_frame = nk / max(1, n_frames)
        h2d = 2 * W * H * F
        d2h = int(kp_per_frame * 60 + 16) * F
        cpu_clients = max(1, (os.cpu_count() or 2) // 2)  # all host cores, two threads per client like the reference
        cpu_fps, cpu_dt, cpu_kind = cpu_extract_fps(max(2, args.sample_frames // cpu_clients), clients=cpu_clients)
        cpu_one_fps = cpu_extract_fps(min(32, args.sample_frames), clients=1)[0]
        cpu_fps_frame, _, _ = cpu_extract_fps(max(8, args.sample_frames // 4), clients=1, stereo=True)
        lat.sort()
        line = {
            "metric": "stereo_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": bench_config(),
            "us_per_frame": {"throughput": 1e3 * dev_ms / n_frames, "step_median": 1e3 * float(np.median(step_ms)) / F,
                             "step_p95": 1e3 * float(np.percentile(step_ms, 95)) / F,
                             "latency_median": lat[len(lat) // 2], "latency_p95": lat[int(len(lat) * 0.95)],
                             "note": "throughput = " + str(IN_FLIGHT) + " frames in flight; latency = one frame alone on the GPU, CUDA events per frame"},
            "e2e": {"value": world * n_frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps, "us_per_frame": 1e3 * e2e_ms / n_frames,
                    "note": "corb_orb_extract_pair_submit / _wait on " + str(IN_FLIGHT) + " handle pairs from one thread: page-locked host images in (copy-engine memcpy nodes, corb_orb_set_host_transfer 1), "
                            "host keypoints + descriptors out (read in place from the handles' page-locked result buffers)"},
            "e2e_one_in_flight": {"value": world * n_frames / (e2e_block_ms * 1e-3), "unit": "frames/s",
                                  "us_per_frame": 1e3 * e2e_block_ms / n_frames,
                                  "note": "the blocking corb_orb_extract_pair (images read in place by the import kernel)"},
            "e2e_stereo_frame": {"value": world * n_frames / (e2e_frame_ms * 1e-3), "unit": "frames/s",
                                 "us_per_frame": 1e3 * e2e_frame_ms / n_frames, "stereo_matches_per_frame": n_stereo / max(1, n_frames),
                                 "note": "corb_frame_stereo_submit / _wait: ExtractORB left+right and Frame::ComputeStereoMatches on the "
                                         "GPU, " + str(IN_FLIGHT) + " frames in flight; the pyramids never leave HBM"},
            "gpu_launches": exl.launches_per_extract() * n_frames,  # one set of launches per stereo pair (grid z = 2)
            "tma": exl.uses_tma(),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "scope": "whole stereo frame (every kernel of the path)", "achieved": achieved,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                         "peak_source": pk_kind, "algorithmic_bytes_per_frame": 2 * ALGO_BYTES_PER_IMAGE,
                         "how": "achieved = 4 054 364 B (SURVEY.md section 8d) x frames/s per GPU; traffic = dram read + write summed over "
                                "the launches of one stereo frame in the committed ncu --set full capture",
                         "traffic_source": cap_src, "dominant_kernel": dom, "kernels": kernels},
            "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": 2 * cpu_clients, "kind": cpu_kind, "cpu_model": cpu_model(),
                             "one_client_value": cpu_one_fps,
                             "sample": "%d stereo frames over %d concurrent clients, %s, L/R on two threads each (Frame.cc:78-81), %.1f s"
                                       % (max(2, args.sample_frames // cpu_clients) * cpu_clients, cpu_clients, "the reference's own ORBextractor.cc (oracle/_ref/libref.so)"
                                          if cpu_kind == "reference" else "oracle port", cpu_dt),
                             "host_cores_available": os.cpu_count(), "stereo_frame_value": cpu_fps_frame,
                             "stereo_frame_kind": "port (Frame::ComputeStereoMatches needs the oracle's pyramids)"},
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
                                       "%d LM iterations, bRobust=false, seed 7; edge arrays in page-locked host memory" % (BA_P, BA_L, ba_edges, BA_ITERS),
                           "sharding": "landmarks l %% N per rank, poses replicated, all-reduce(sum, fp64) of the reduced "
                                       "camera system per LM trial" if world > 1 else "single GPU, no collective"},
                "ms_total": ba_ms, "ms_in_library": ba_info["ms_total"], "ms_reduced_camera_solves": ba_info["ms_solve"],
                "ms_setup": ba_info["ms_setup"],
                "iterations": ba_info["iterations"], "trials": ba_info["n_trials"], "trial_accepted": ba_info["trial_accepted"],
                "chi2_initial": ba_info["chi2_initial"], "chi2_final": ba_info["chi2_final"], "rms_px": ba_rms,
                "reduced_blocks": ba_info["reduced_blocks"], "band_chunks": ba_info["band_chunks"], "border_poses": ba_info["border_poses"],
                "parity_vs_single": ba_parity, "server_path_one_process": ba_server,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                             "algorithmic_bytes_per_iteration": ba_bytes // max(1, ba_info["iterations"]),
                             "note": "whole-call wall time incl. host structure build and H2D/D2H; the reduced-camera solve "
                                     "(chunked block-skyline Cholesky) is bound by the serial latency of a block column"},
                "cpu_baseline": dict(cpu_ba(BA_P, BA_L), cpu_model=cpu_model()) if not args.no_cpu_ba else None,
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
    for a_, b_ in pairs:
        a_.close(); b_.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="timed steps of %d stereo frames each" % FRAMES_PER_STEP)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sample-frames", type=int, default=128, help="stereo frames of the bounded CPU baseline sample")
    ap.add_argument("--no-ba", action="store_true", help="skip the global-BA half of the metric")
    ap.add_argument("--no-cpu-ba", action="store_true", help="skip the CPU BA baseline (about 10 s)")
    ap.add_argument("--no-matcher", action="store_true", help="skip the SearchByBoW / vocabulary / score section")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
