"""Timeline of one frame graph from %globaltimer stamps (library built with `make EXTRA=-DCORB_TIMELINE`)."""
import sys, ctypes as C; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from corb_slam_b200 import ORBextractor, extract_stereo_device, _lib
from corb_slam_b200.synth import stereo_frame, frame_seed
L = _lib.lib()
L.corb_debug_timeline.argtypes = [C.c_void_p, C.c_int]
P = (2000, 1.2, 8, 20, 7)
pair = len(sys.argv) > 1 and sys.argv[1] in ("pair", "pair_host")
host = len(sys.argv) > 1 and sys.argv[1] == "pair_host"
from corb_slam_b200 import extract_stereo
import time
exl, exr = ORBextractor(*P), ORBextractor(*P)
exl.copy_outputs = exr.copy_outputs = False
l, r = stereo_frame(frame_seed(0))
pl, pr = torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()
dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
names = {0: "import", 48: "blur", 49: "orient_desc"}
for i in range(1, 8): names[i] = "resize[%d]" % i
for i in range(8): names[16 + i] = "fast[%d]" % i; names[32 + i] = "octtree[%d]" % i
for rep in range(6):
    L.corb_debug_timeline(None, 1)
    torch.cuda.synchronize()
    if host:
        t_host0 = time.perf_counter()
        extract_stereo(exl, exr, pl.numpy(), pr.numpy())
        t_host1 = time.perf_counter()
    elif pair:
        extract_stereo_device(exl, exr, dl.data_ptr(), dr.data_ptr(), 1242, 375, 1242); exl.sync(); exr.sync()
    else:
        exl.extract_device(dl.data_ptr(), 1242, 375, 1242); exl.sync()
    out = np.zeros((2, 64), np.uint64)
    L.corb_debug_timeline(out.ctypes.data, 0)
    if rep < 3: continue
    t0 = min(int(out[0][k]) for k in names if out[1][k] > 0)
    if host: print("host wall time of the call: %.1f us" % ((t_host1 - t_host0) * 1e6))
    print("--- rep %d (%s), times in us from the first kernel start" % (rep, "stereo pair" if pair else "one image"))
    for k in sorted(names, key=lambda k: int(out[0][k])):
        if out[1][k] == 0: continue
        a, b = (int(out[0][k]) - t0) / 1e3, (int(out[1][k]) - t0) / 1e3
        print("%-14s %7.2f -> %7.2f  (%5.2f)  %s" % (names[k], a, b, b - a, " " * int(a / 1.5) + "#" * max(1, int((b - a) / 1.5))))
