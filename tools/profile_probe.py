"""Small driver for ncu captures: stereo frames through corb_frame_stereo, one SearchByBoW batch + vocabulary
transform + L1 scores, and one global BA (P=500, L=50 000)."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from corb_slam_b200 import ORBextractor, frame_stereo, ORBmatcher, ORBVocabulary, BowFeatures, Optimizer
from corb_slam_b200.synth import stereo_frame, ba_problem

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "orb"):
    exl, exr = ORBextractor(2000, 1.2, 8, 20, 7), ORBextractor(2000, 1.2, 8, 20, 7)
    for i in range(3):
        l, r = stereo_frame(1234 + i)
        (kl, dl), (kr, dr), ur, dp = frame_stereo(exl, exr, l, r, 386.1448, 386.1448 / 718.856)
    print("frame", len(kl), len(kr), int((ur >= 0).sum()), "tma", exl.uses_tma())
if what in ("all", "match"):
    sys.path.insert(0, '/root/repo')
    import bench
    print(bench.bench_matcher(0, with_cpu=False)["search_by_bow"])
if what in ("all", "ba"):
    out, info = Optimizer.BundleAdjustment(ba_problem(500, 50000, seed=7), 3, bRobust=False)
    print("ba", info["iterations"], info["chi2_final"], info["ms_total"], info["ms_solve"])
