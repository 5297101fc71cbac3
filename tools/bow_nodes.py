"""FeatureVector node sizes of a bench frame on the real vocabulary (levelsup 4): python tools/bow_nodes.py"""
import sys; sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np
import bench
from corb_slam_b200 import ORBVocabulary
(qk, base), cands = bench.matcher_workload(0, n_cand=2)
v = ORBVocabulary(device=0); v.loadFromTextFile(bench.vocabulary_text())
for d in (base, cands[0][1]):
    bw, bv, fn, fo, fi = v.transform(d, 4)
    sz = np.diff(fo)
    print("features", len(d), "words", len(bw), "nodes", len(fn), "node size max", sz.max(), "p50/p90/p99", np.percentile(sz, [50, 90, 99]), "over 128:", int((sz > 128).sum()))
