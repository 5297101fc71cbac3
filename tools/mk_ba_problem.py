import sys; sys.path.insert(0,'/root/repo')
from corb_slam_b200.synth import ba_problem
from corb_slam_b200.ba_file import write_problem
write_problem('/tmp/p.bin', ba_problem(2000, 200000, seed=7))
