"""Times the global BA on the bench problem (no oracle): python tools/ba_time.py [P L]"""
import sys; sys.path.insert(0, '/root/repo')
from corb_slam_b200 import Optimizer
from corb_slam_b200.synth import ba_problem
P, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 200000)
prob = ba_problem(P, L, seed=7)
Optimizer.BundleAdjustment(ba_problem(50, 2000, seed=1), 2, bRobust=False)
Optimizer.BundleAdjustment(ba_problem(P, L, seed=7), 1, bRobust=False)  # sizes the per-device arena, like bench.py
out, g = Optimizer.BundleAdjustment(prob, 10, bRobust=False)
print("setup %.1f ms, chunks %d, separators %d" % (g["ms_setup"], g["band_chunks"], g["separator_poses"]))
print("P=%d L=%d: ms_total %.1f ms_solve %.1f iterations %d trials %d chi2 %.6f blocks %d border %d max_active %d" % (
    P, L, g["ms_total"], g["ms_solve"], g["iterations"], g["n_trials"], g["chi2_final"], g["reduced_blocks"], g["border_poses"], g["max_active_rows"]))
