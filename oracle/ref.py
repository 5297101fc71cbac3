"""ctypes bindings of oracle/_ref/libref.so = the REFERENCE'S OWN SOURCES (ORBextractor.cc, ORBmatcher.cc, Frame.cc,
PnPsolver.cc, DBoW2) compiled unmodified against the stub OpenCV of oracle/refbuild (TEST INFRASTRUCTURE ONLY).

libref.so is built in this container from /root/reference by oracle/refbuild/Makefile and travels to the GPU box as a
prebuilt file (oracle/_ref/ is git-ignored, not gpurun-ignored). Only tests/, smoke() and bench.py's reference legs
may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import KP_DTYPE, _p, _u8p, _i32p, _f32p, _f64p, build as _build_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE = "/root/reference/corbslam_client"
_lib = None


def available():
    return os.path.exists(LIB_PATH) or os.path.isdir(REFERENCE)


def build(force=False):
    """Compile the reference sources where they lie (only possible where /root/reference exists)."""
    _build_oracle()
    if os.path.isdir(REFERENCE):
        subprocess.run(["make", "-C", os.path.join(_HERE, "refbuild"), "-s", "-j8"] + (["-B"] if force else []), check=True)
    elif not os.path.exists(LIB_PATH):
        raise RuntimeError("oracle/_ref/libref.so is missing and /root/reference is not mounted")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.ref_orb_create.restype = C.c_void_p
        L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_orb_destroy.argtypes = [C.c_void_p]
        L.ref_orb_extract.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p, C.c_int]
        L.ref_orb_levels.argtypes = [C.c_void_p]
        L.ref_orb_tables.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p]
        L.ref_orb_pyramid.restype = C.POINTER(C.c_uint8)
        L.ref_orb_pyramid.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _i32p]
        _lib = L
    return _lib


class ORBextractor:
    """ORB_SLAM2::ORBextractor itself (corbslam_client/src/ORBextractor.cc, unmodified)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self._h = lib().ref_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels = nfeatures, nlevels
        n = nlevels
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = (np.empty(n, np.float32) for _ in range(4))
        lib().ref_orb_tables(self._h, _p(self.scale, _f32p), _p(self.inv_scale, _f32p), _p(self.sigma2, _f32p),
                             _p(self.inv_sigma2, _f32p))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_orb_destroy(self._h)
            self._h = None

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures * 2 + 64 * self.nlevels
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().ref_orb_extract(self._h, _p(img, _u8p), w, h, img.strides[0], kps.ctypes.data_as(C.c_void_p), _p(desc, _u8p), cap)
        if n < 0:
            raise RuntimeError("capacity %d < %d keypoints" % (cap, -n))
        return kps[:n].copy(), desc[:n].copy()

    def pyramid(self, level):
        """mvImagePyramid[level] after the last call."""
        w, h, s = C.c_int32(), C.c_int32(), C.c_int32()
        p = lib().ref_orb_pyramid(self._h, level, C.byref(w), C.byref(h), C.byref(s))
        a = np.ctypeslib.as_array(p, shape=(h.value, s.value))
        return a[:, :w.value].copy()
