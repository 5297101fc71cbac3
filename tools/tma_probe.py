import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from corb_slam_b200 import ORBextractor
from corb_slam_b200.synth import stereo_frame
left, _ = stereo_frame(1234)
ex = ORBextractor(2000, 1.2, 8, 20, 7)
k, d = ex(left)
print("uses_tma", ex.uses_tma(), len(k))
