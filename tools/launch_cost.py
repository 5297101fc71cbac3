"""Host-side cost of one stereo-frame launch (corb_orb_extract_pair_device: patch the import node + cudaGraphLaunch) against
the device time per frame, with several frames in flight. Usage: python tools/launch_cost.py [in_flight]"""
import sys, time
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np, torch
from corb_slam_b200 import ORBextractor, extract_stereo_device
from corb_slam_b200.synth import stereo_frame, frame_seed
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 2
P = (2000, 1.2, 8, 20, 7)
pairs = [(ORBextractor(*P), ORBextractor(*P)) for _ in range(NP)]
fr = [stereo_frame(frame_seed(i)) for i in range(8)]
dev = [(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()) for l, r in fr]
W, H = 1242, 375
for mode in ("distinct images (import node patched every launch)", "same image (no patch)"):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in range(256):
            i = f % 8 if mode.startswith("distinct") else 0
            extract_stereo_device(*pairs[f % NP], dev[i][0].data_ptr(), dev[i][1].data_ptr(), W, H, W)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print("%d in flight, %s: host enqueue %.1f us/frame, total %.1f us/frame" % (NP, mode, (t1 - t0) / 256 * 1e6, (t2 - t0) / 256 * 1e6))
