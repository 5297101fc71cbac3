"""Times both SearchByProjection variants on the synthetic scene (GPU through the C ABI vs the oracle on one host thread)."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
import oracle
from oracle import _proj_bind as PB
from corb_slam_b200 import FrameView, ORBmatcher
from corb_slam_b200.synth import projection_scene
oracle.lib()
sc = projection_scene(1)
c = sc["cur"]
fv = FrameView(c["x"], c["y"], c["octave"], c["angle"], c["desc"], c["u_right"], sc["scales"], sc["bounds"], sc["K"], sc["mbf"], sc["Tcw"], taken=sc["taken"])
m = ORBmatcher(0.8, True)
la = (fv, sc["last_valid"], sc["Xw"], sc["mp_desc"], sc["last"]["octave"], sc["last"]["angle"], sc["Tlw"], 15.0)
ma = (fv, sc["in_view"], sc["proj"], sc["level"], sc["view_cos"], sc["mp_desc"])
for _ in range(3):
    m.SearchByProjectionLastFrame(*la, last_blocks=sc["last_blocks"]); m.SearchByProjectionMapPoints(*ma)
R = 200
t0 = time.perf_counter()
for _ in range(R): m.SearchByProjectionLastFrame(*la, last_blocks=sc["last_blocks"])
t1 = time.perf_counter()
for _ in range(R): m.SearchByProjectionMapPoints(*ma)
t2 = time.perf_counter()
cs = fv.c_struct()
for _ in range(20): PB.search_by_projection_last(cs, fv.n, sc["last_valid"], sc["last_blocks"], sc["Xw"], sc["mp_desc"], sc["last"]["octave"], sc["last"]["angle"], sc["Tlw"], 15.0, False, True)
t3 = time.perf_counter()
for _ in range(20): PB.search_by_projection_map(cs, fv.n, sc["in_view"], None, sc["proj"], sc["level"], sc["view_cos"], sc["mp_desc"], 1.0, 0.8)
t4 = time.perf_counter()
print("GPU last-frame %.3f ms  map-points %.3f ms | CPU oracle last-frame %.3f ms  map-points %.3f ms" % (
    1e3 * (t1 - t0) / R, 1e3 * (t2 - t1) / R, 1e3 * (t3 - t2) / 20, 1e3 * (t4 - t3) / 20))
