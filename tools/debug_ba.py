import sys; sys.path.insert(0, '/root/repo')
import numpy as np, oracle
from oracle import _ba_bind as B
from corb_slam_b200 import Optimizer
from corb_slam_b200.synth import ba_problem
oracle.lib()
P, L = int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 2000
prob = ba_problem(P, L, seed=7, n_fusion=20)
o_out, o = B.solve(prob, 10)
g_out, g = Optimizer.BundleAdjustment(prob, 10, bRobust=False)
for k in ("iterations", "n_trials", "chi2_initial", "chi2_final", "lambda_initial", "lambda_final", "solver_failures", "trial_accepted"):
    print(k, o[k], g[k])
print("oracle chi2", o["trial_chi2"])
print("gpu    chi2", g["trial_chi2"])
print("ms", g["ms_total"], g["ms_solve"], g["reduced_blocks"], g["border_poses"], g["max_active_rows"])
print("dX", np.abs(g_out["point_xyz"] - o_out["point_xyz"]).max(), "dt", np.abs(g_out["pose_t"] - o_out["pose_t"]).max())
