"""Per-kernel event profile of one image + bench-style device/e2e timing (quick A/B of extractor changes)."""
import sys, os, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from corb_slam_b200 import ORBextractor, extract_stereo, extract_stereo_device
from corb_slam_b200.synth import stereo_frame, frame_seed
P = (2000, 1.2, 8, 20, 7)
exl, exr = ORBextractor(*P), ORBextractor(*P)
exl.copy_outputs = exr.copy_outputs = False
frames = [stereo_frame(frame_seed(i)) for i in range(4)]
pinned = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) for l, r in frames]
dev = [(a.cuda(), b.cuda()) for a, b in pinned]
npin = [(a.numpy(), b.numpy()) for a, b in pinned]
for i in range(5):
    extract_stereo_device(exl, exr, dev[i % 4][0].data_ptr(), dev[i % 4][1].data_ptr(), 1242, 375, 1242); exl.sync(); exr.sync()
    extract_stereo(exl, exr, npin[i % 4][0], npin[i % 4][1])
n = 300
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(n):
    extract_stereo_device(exl, exr, dev[i % 4][0].data_ptr(), dev[i % 4][1].data_ptr(), 1242, 375, 1242); exl.sync(); exr.sync()
t1 = time.perf_counter()
for i in range(n):
    extract_stereo(exl, exr, npin[i % 4][0], npin[i % 4][1])
t2 = time.perf_counter()
print("device-resident us/frame %.1f   e2e us/frame %.1f" % ((t1 - t0) / n * 1e6, (t2 - t1) / n * 1e6))
exl.extract_device(dev[0][0].data_ptr(), 1242, 375, 1242); exl.sync()
for name, ms in exl.profile(reps=20):
    print("%-18s %.2f us" % (name, ms * 1e3))
