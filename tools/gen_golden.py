"""Generates tests/golden/opencv_primitives.npz: outputs of the real OpenCV (cv2 4.13.0 in the build container) for the
primitives the reference's ORB front end calls: FAST-9/16 with NMS, resize INTER_LINEAR u8, GaussianBlur
7x7 sigma 2 REFLECT_101 u8, fastAtan2. The oracle's closed-form models are pinned against these vectors by
tests/test_oracle_cpu.py (the GPU box needs neither cv2 nor /root/reference).

    python tools/gen_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from corb_slam_b200.synth import stereo_frame  # noqa: E402


def main():
    rng = np.random.default_rng(20261017)
    out = {"cv2_version": np.array(cv2.__version__)}
    left, _ = stereo_frame(42, w=320, h=200)
    noise = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    out["img_scene"] = left
    out["img_noise"] = noise
    # resize chain with the reference's level sizes (cvRound(w / 1.2^l)) and two odd sizes
    for name, img, sizes in (("scene", left, [(267, 167), (222, 139), (185, 116)]), ("noise", noise, [(109, 81), (91, 67), (40, 23)])):
        cur = img
        for i, (w, h) in enumerate(sizes):
            cur = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
            out["resize_%s_%d" % (name, i)] = cur
    out["blur_scene"] = cv2.GaussianBlur(left, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    out["blur_noise"] = cv2.GaussianBlur(noise, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    tiny = noise[:9, :11].copy()
    out["img_tiny"] = tiny
    out["blur_tiny"] = cv2.GaussianBlur(tiny, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    for name, img in (("scene", left), ("noise", noise), ("cell", np.ascontiguousarray(left[30:68, 100:137]))):
        for th in (20, 7):
            det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kps = det.detect(img)
            out["fast_%s_%d" % (name, th)] = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
    out["img_cell"] = np.ascontiguousarray(left[30:68, 100:137])
    ys = np.concatenate([rng.integers(-40000, 40000, 500), [0, 0, 5, -5, 1, -1, 0]]).astype(np.float32)
    xs = np.concatenate([rng.integers(-40000, 40000, 500), [0, 7, 0, 0, 1, -1, -3]]).astype(np.float32)
    out["atan2_y"], out["atan2_x"] = ys, xs
    out["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    # cv::Mat products of the projection matchers (ORBmatcher.cc:1483,1488,1504) = cv::gemm on CV_32F:
    # Rcw*x+tcw, -Rcw.t()*t, with poses / points of realistic magnitude
    n = 400
    ang = rng.uniform(-0.3, 0.3, (n, 3))
    Rs = np.empty((n, 3, 3), np.float32)
    for i, (a, b, c) in enumerate(ang):
        Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
        Rs[i] = (Rz @ Ry @ Rx).astype(np.float32)
    ts = rng.uniform(-50, 50, (n, 3, 1)).astype(np.float32)
    xs3 = rng.uniform(-80, 80, (n, 3, 1)).astype(np.float32)
    out["gemm_R"], out["gemm_t"], out["gemm_x"] = Rs, ts, xs3
    out["gemm_Rx_plus_t"] = np.stack([cv2.gemm(Rs[i], xs3[i], 1.0, ts[i], 1.0) for i in range(n)])
    out["gemm_minus_Rt_t"] = np.stack([cv2.gemm(Rs[i], ts[i], -1.0, None, 0.0, flags=cv2.GEMM_1_T) for i in range(n)])
    path = os.path.join(ROOT, "tests", "golden", "opencv_primitives.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.startswith("fast")})


if __name__ == "__main__":
    main()


def oracle_regression():
    """tests/golden/oracle_extract_640x240.npz: the oracle's own output on one seeded frame. Not a reference vector —
    it only guards the composed oracle (cell grid, quadtree order, epilogue) against accidental change."""
    import oracle
    left, _ = stereo_frame(1234, w=640, h=240)
    ex = oracle.OrbExtractor(500, 1.2, 4, 20, 7)
    kps, desc = ex(left)
    path = os.path.join(ROOT, "tests", "golden", "oracle_extract_640x240.npz")
    np.savez_compressed(path, kps=kps, desc=desc, level_count=np.array([ex.level_count(l) for l in range(4)]),
                        cand_count=np.array([len(ex.candidates(l)) for l in range(4)]))
    print("wrote", path, len(kps), "keypoints")


if __name__ == "__main__" and "--oracle" in sys.argv:
    oracle_regression()
