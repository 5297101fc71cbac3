"""The C-ABI library loads without a GPU, exports every symbol include/corb_b200.h declares, and its compute entry
points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "corb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"CORB_API\s+[\w\s\*]+?\b(corb_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from corb_slam_b200 import _lib
    names = _declared()
    assert len(names) >= 35, names
    L = _lib.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(names) <= exported
    assert all(n.startswith("corb_") for n in exported if not n.startswith("_")), exported  # nothing else leaks


def test_library_was_built_for_sm_100a_only():
    from corb_slam_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "corb_slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt and "oracle/" not in txt.replace("oracle/orb_oracle.cpp", "").replace("oracle/match_oracle.cpp", ""), f


def test_no_cpu_fallback_without_a_gpu():
    from corb_slam_b200 import CorbError, ORBextractor, ORBmatcher, Optimizer, _lib
    if _lib.lib().corb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(CorbError):
        ORBextractor(2000, 1.2, 8, 20, 7)
    with pytest.raises(CorbError):
        ORBmatcher(0.75, True)
    from corb_slam_b200.synth import ba_problem
    with pytest.raises(CorbError):
        Optimizer.BundleAdjustment(ba_problem(5, 50, seed=0), 1, bRobust=False)
    assert b"CUDA" in _lib.lib().corb_last_error() or b"cuda" in _lib.lib().corb_last_error()
    from corb_slam_b200 import PnPsolver
    from corb_slam_b200.synth import pnp_problem
    p = pnp_problem(0, n=40)
    s = PnPsolver(p["p2d"], np.zeros(40, int), [1.0], p["K"], p["p3d"], np.ones(40, bool))
    with pytest.raises(CorbError):
        s.iterate(5)


def test_argument_validation_without_a_gpu():
    from corb_slam_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.corb_orb_create(0, 1.2, 8, 20, 7, 0, C.byref(h)) == _lib.ERR_INVALID     # nfeatures < 1
    assert L.corb_orb_create(2000, 1.0, 8, 20, 7, 0, C.byref(h)) == _lib.ERR_INVALID  # scale factor must exceed 1
    assert L.corb_orb_create(2000, 1.2, 99, 20, 7, 0, C.byref(h)) == _lib.ERR_INVALID
    assert L.corb_orb_levels(None) == 0 and L.corb_orb_capacity(None, 100, 100) < 0
    assert L.corb_version() >= 100
