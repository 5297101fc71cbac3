"""Generates tests/golden/pnp_cv2.npz: outputs of the real OpenCV (cv2 4.13.0 in the build container) for the arithmetic
PnPsolver delegates to it: cvSVD (cv2.SVDecomp), cvSolve(CV_SVD) (cv2.solve DECOMP_SVD),
cvInvert(CV_SVD) (cv2.invert DECOMP_SVD), and OpenCV's own copy of the EPnP code the reference vendors in
PnPsolver.cc:420-962 (cv2.solvePnP, SOLVEPNP_EPNP). tests/test_pnp_cpu.py pins oracle/pnp_oracle.cpp against these
vectors (the GPU box needs neither cv2 nor /root/reference).

    python tools/gen_golden_pnp.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from corb_slam_b200.synth import pnp_problem  # noqa: E402


def main():
    rng = np.random.default_rng(20261017)
    out = {"cv2_version": np.array(cv2.__version__)}
    shapes = [(12, 12), (3, 3), (6, 4), (6, 3), (6, 5)]
    for si, (m, n) in enumerate(shapes):
        for rep in range(4):
            A = rng.standard_normal((m, n))
            if m == 12:  # symmetric positive definite like M^T M (well separated spectrum)
                B = rng.standard_normal((12 + 8 * rep, 12))
                A = B.T @ B
            w, u, vt = cv2.SVDecomp(A)
            key = "svd_%d_%d" % (si, rep)
            out[key + "_A"], out[key + "_w"], out[key + "_u"], out[key + "_vt"] = A, w.ravel(), u, vt
            if m <= 6:
                b = rng.standard_normal(m)
                ok, x = cv2.solve(A, b.reshape(-1, 1), flags=cv2.DECOMP_SVD)
                out[key + "_b"], out[key + "_x"] = b, x.ravel()
            if m == 3:
                out[key + "_inv"] = cv2.invert(A, flags=cv2.DECOMP_SVD)[1]
    # EPnP on 6..200 correspondences, inliers only, with and without pixel noise
    ci = 0
    for n in (6, 8, 15, 60, 200):
        for noise in (0.0, 0.7):
            p = pnp_problem(1000 + ci, n=n, outlier_fraction=0.0, pixel_noise=noise)
            fx, fy, cx, cy = [float(v) for v in p["K"]]
            K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
            Xw, uv = p["p3d"].astype(np.float64), p["p2d"].astype(np.float64)
            ok, rvec, tvec = cv2.solvePnP(Xw, uv, K, None, flags=cv2.SOLVEPNP_EPNP)
            assert ok
            key = "epnp_%d" % ci
            out[key + "_Xw"], out[key + "_uv"], out[key + "_K"] = Xw, uv, np.array([fx, fy, cx, cy])
            out[key + "_R"], out[key + "_t"] = cv2.Rodrigues(rvec)[0], tvec.ravel()
            out[key + "_Rtrue"], out[key + "_ttrue"] = p["Tcw"][:, :3], p["Tcw"][:, 3]
            ci += 1
    out["n_epnp"] = np.array(ci)
    path = os.path.join(ROOT, "tests", "golden", "pnp_cv2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
