"""ctypes signatures + numpy wrappers of oracle/proj_oracle.cpp (TEST INFRASTRUCTURE ONLY). The frame view struct has the
layout of corb_frame_view, so tests build it once (corb_slam_b200.frame.FrameView.c_struct) and hand it to both sides."""
import ctypes as C

import numpy as np

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_L = None


def bind(L):
    global _L
    _L = L
    vp = C.c_void_p
    L.oracle_gemm3.argtypes = [_f32p, C.c_int, C.c_double, _f32p, C.c_double, _f32p, _f32p]
    L.oracle_gemm3.restype = None
    L.oracle_search_by_projection_last.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp]
    L.oracle_search_by_projection_map.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp]


def _lib():
    from . import lib
    lib()
    return _L


def gemm3(M34, transpose, alpha, v, beta=0.0, c=None):
    M = np.ascontiguousarray(M34, np.float32).reshape(-1)
    v = np.ascontiguousarray(v, np.float32)
    out = np.empty(3, np.float32)
    cc = None if c is None else np.ascontiguousarray(c, np.float32)
    _lib().oracle_gemm3(M.ctypes.data_as(_f32p), int(transpose), float(alpha), v.ctypes.data_as(_f32p), float(beta),
                        cc.ctypes.data_as(_f32p) if cc is not None else None, out.ctypes.data_as(_f32p))
    return out


def _ptr(a):
    return a.ctypes.data if a is not None else None


def search_by_projection_last(view_struct, n_cur, last_valid, last_blocks, last_xyz, last_mp_desc, last_octave, last_angle, Tlw, th,
                              mono, check_ori):
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    valid, blocks = u8(last_valid), u8(last_blocks)
    xyz = np.ascontiguousarray(last_xyz, np.float32); desc = np.ascontiguousarray(last_mp_desc, np.uint8)
    octv = np.ascontiguousarray(last_octave, np.int32); ang = np.ascontiguousarray(last_angle, np.float32)
    T = np.ascontiguousarray(np.asarray(Tlw, np.float32).reshape(-1)[:12])
    match = np.full(n_cur, -1, np.int32)
    n = _lib().oracle_search_by_projection_last(C.addressof(view_struct), len(valid), _ptr(valid), _ptr(blocks), _ptr(xyz), _ptr(desc),
                                                _ptr(octv), _ptr(ang), _ptr(T), float(th), int(mono), int(check_ori), _ptr(match))
    return match, n


def search_by_projection_map(view_struct, n_frame, in_view, blocks, proj, level, view_cos, mp_desc, th, nnratio):
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    iv, bl = u8(in_view), u8(blocks)
    pr = np.ascontiguousarray(proj, np.float32); lv = np.ascontiguousarray(level, np.int32)
    vc = np.ascontiguousarray(view_cos, np.float32); desc = np.ascontiguousarray(mp_desc, np.uint8)
    match = np.full(n_frame, -1, np.int32)
    n = _lib().oracle_search_by_projection_map(C.addressof(view_struct), len(iv), _ptr(iv), _ptr(bl), _ptr(pr), _ptr(lv), _ptr(vc),
                                               _ptr(desc), float(th), float(nnratio), _ptr(match))
    return match, n
