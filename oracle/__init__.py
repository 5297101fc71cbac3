"""ctypes bindings of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package. The product path (corb_slam_b200 -> libcorb_b200.so) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile the C++ restatement (gcc only, no dependencies)."""
    srcs = [f for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in srcs)
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def _p(a, t):
    return a.ctypes.data_as(t)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_orb_create.restype = C.c_void_p
        L.oracle_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.oracle_orb_destroy.argtypes = [C.c_void_p]
        L.oracle_orb_levels.argtypes = [C.c_void_p]
        L.oracle_orb_tables.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _i32p, _i32p]
        L.oracle_orb_level_size.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, _i32p]
        L.oracle_orb_capacity.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_orb_extract.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p]
        L.oracle_orb_pyramid.restype = _u8p
        L.oracle_orb_pyramid.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p]
        L.oracle_orb_blurred.restype = _u8p
        L.oracle_orb_blurred.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p]
        L.oracle_orb_candidates.argtypes = [C.c_void_p, C.c_int, C.POINTER(_i32p)]
        L.oracle_orb_level_count.argtypes = [C.c_void_p, C.c_int]
        L.oracle_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _u8p, C.c_int, C.c_void_p, _u8p, C.c_int, C.c_float,
                                            C.c_float, _f32p, _f32p]
        L.oracle_resize_linear_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]
        L.oracle_gaussian7_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.oracle_fast_score.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.oracle_fast_detect.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int]
        L.oracle_fast_atan2.restype = C.c_float
        L.oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.oracle_cv_round_f.argtypes = [C.c_float]
        L.oracle_ic_angle.restype = C.c_float
        L.oracle_ic_angle.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p]
        L.oracle_brief.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, _u8p]
        _bind_match(L)
        _bind_ba(L)
        if hasattr(L, "oracle_search_by_projection_last"):
            from . import _proj_bind
            _proj_bind.bind(L)
        if hasattr(L, "oracle_pnp_iterate"):
            from . import _pnp_bind
            _pnp_bind.bind(L)
        _lib = L
    return _lib


def _bind_match(L):
    if not hasattr(L, "oracle_hamming256"):
        return
    from . import _match_bind
    _match_bind.bind(L)


def _bind_ba(L):
    if not hasattr(L, "oracle_ba_solve"):
        return
    from . import _ba_bind
    _ba_bind.bind(L)


# ----------------------------------------------------------------------------- primitives
def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_linear_u8(_p(src, _u8p), src.shape[1], src.shape[0], src.shape[1], _p(dst, _u8p), dw, dh, dw)
    return dst


def gaussian7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().oracle_gaussian7_u8(_p(src, _u8p), src.shape[1], src.shape[0], src.shape[1], _p(dst, _u8p), src.shape[1])
    return dst


def fast_score(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().oracle_fast_score(_p(src, _u8p), src.shape[1], src.shape[0], src.shape[1], _p(dst, _u8p), src.shape[1])
    return dst


def fast_detect(src, th):
    src = np.ascontiguousarray(src, np.uint8)
    cap = src.size // 4 + 16
    out = np.empty((cap, 3), np.int32)
    n = lib().oracle_fast_detect(_p(src, _u8p), src.shape[1], src.shape[0], src.shape[1], th, _p(out, _i32p), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return lib().oracle_fast_atan2(float(y), float(x))


def cv_round(v):
    return lib().oracle_cv_round_f(float(v))


def ic_angle(img, x, y, umax):
    img = np.ascontiguousarray(img, np.uint8)
    um = np.ascontiguousarray(umax, np.int32)
    return lib().oracle_ic_angle(_p(img, _u8p), img.shape[1], int(x), int(y), _p(um, _i32p))


def brief(img, x, y, angle_deg):
    img = np.ascontiguousarray(img, np.uint8)
    d = np.empty(32, np.uint8)
    lib().oracle_brief(_p(img, _u8p), img.shape[1], int(x), int(y), float(angle_deg), _p(d, _u8p))
    return d


# ----------------------------------------------------------------------------- extractor
class OrbExtractor:
    """Mirror of ORB_SLAM2::ORBextractor (corbslam_client/include/ORBextractor.h:45-112) on the CPU oracle."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self._h = lib().oracle_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        if not self._h:
            raise ValueError("invalid ORB parameters")
        self.nlevels = nlevels
        n = nlevels
        self.scale = np.empty(n, np.float32)
        self.inv_scale = np.empty(n, np.float32)
        self.sigma2 = np.empty(n, np.float32)
        self.inv_sigma2 = np.empty(n, np.float32)
        self.quota = np.empty(n, np.int32)
        self.umax = np.empty(16, np.int32)
        lib().oracle_orb_tables(self._h, _p(self.scale, _f32p), _p(self.inv_scale, _f32p), _p(self.sigma2, _f32p),
                                _p(self.inv_sigma2, _f32p), _p(self.quota, _i32p), _p(self.umax, _i32p))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_orb_destroy(self._h)
            self._h = None

    def level_size(self, level, w, h):
        lw, lh = C.c_int32(), C.c_int32()
        lib().oracle_orb_level_size(self._h, level, w, h, C.byref(lw), C.byref(lh))
        return lw.value, lh.value

    def capacity(self, w, h):
        return lib().oracle_orb_capacity(self._h, w, h)

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.capacity(w, h)
        if cap < 0:
            raise ValueError("image too small for the configured pyramid")
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().oracle_orb_extract(self._h, _p(img, _u8p), w, h, img.strides[0], kps.ctypes.data_as(C.c_void_p),
                                     _p(desc, _u8p))
        if n < 0:
            raise RuntimeError("oracle_orb_extract failed: %d" % n)
        return kps[:n].copy(), desc[:n].copy()

    def pyramid(self, level):
        w, h = C.c_int32(), C.c_int32()
        p = lib().oracle_orb_pyramid(self._h, level, C.byref(w), C.byref(h))
        return np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()

    def blurred(self, level):
        w, h = C.c_int32(), C.c_int32()
        p = lib().oracle_orb_blurred(self._h, level, C.byref(w), C.byref(h))
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()

    def candidates(self, level):
        ptr = _i32p()
        n = lib().oracle_orb_candidates(self._h, level, C.byref(ptr))
        if n == 0:
            return np.zeros((0, 3), np.int32)
        return np.ctypeslib.as_array(ptr, shape=(n, 3)).copy()

    def level_count(self, level):
        return lib().oracle_orb_level_count(self._h, level)


def stereo_matches(ex_left, ex_right, kps_l, desc_l, kps_r, desc_r, mbf, mb):
    """Frame::ComputeStereoMatches (Frame.cc:470-644) -> (mvuRight, mvDepth, n_kept); needs both extractors' last pyramids."""
    kl = np.ascontiguousarray(kps_l); kr = np.ascontiguousarray(kps_r)
    dl = np.ascontiguousarray(desc_l, np.uint8); dr = np.ascontiguousarray(desc_r, np.uint8)
    ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32)
    n = lib().oracle_stereo_matches(ex_left._h, ex_right._h, kl.ctypes.data_as(C.c_void_p), _p(dl, _u8p), len(kl),
                                    kr.ctypes.data_as(C.c_void_p), _p(dr, _u8p), len(kr), float(mbf), float(mb), _p(ur, _f32p),
                                    _p(dp, _f32p))
    return ur, dp, n
