"""Binary KeyFrame payload (corb_kf_payload_*, include/corb_b200.h): the extractor-produced members of a KeyFrame
(reference: corbslam_client/include/KeyFrame.h:61-87 serialize(), SerializeObject.h:34-61, DataDriver.cc:40-238) as one
version-tagged blob instead of boost text-archive decimals. Host code only."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib

_bound = False


def _L():
    global _bound
    L = lib()
    if not _bound:
        vp = C.c_void_p
        L.corb_kf_payload_bound.restype = C.c_size_t
        L.corb_kf_payload_bound.argtypes = [C.c_int] * 4
        L.corb_kf_payload_encode.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.corb_kf_payload_info.argtypes = [vp, C.c_size_t] + [C.POINTER(C.c_int)] * 5
        L.corb_kf_payload_decode.argtypes = [vp, C.c_size_t] + [vp] * 10
        _bound = True
    return L


def encode(keys, keys_un, u_right, depth, desc, bow, fv):
    """keys / keys_un: KP_DTYPE arrays (keys_un None = identical); bow = (words u32, vals f64); fv = (nodes, off, idx). -> bytes"""
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    un = None if keys_un is None else np.ascontiguousarray(keys_un, KP_DTYPE)
    n = len(keys)
    ur = np.ascontiguousarray(u_right, np.float32); dp = np.ascontiguousarray(depth, np.float32)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32) if n else np.zeros((0, 32), np.uint8)
    bw = np.ascontiguousarray(bow[0], np.uint32); bv = np.ascontiguousarray(bow[1], np.float64)
    fn = np.ascontiguousarray(fv[0], np.uint32); fo = np.ascontiguousarray(fv[1], np.int32); fi = np.ascontiguousarray(fv[2], np.uint32)
    cap = _L().corb_kf_payload_bound(n, len(bw), len(fn), len(fi))
    out = np.empty(cap, np.uint8)
    w = C.c_size_t()
    p = lambda a: a.ctypes.data if a is not None and a.size else None
    check(_L().corb_kf_payload_encode(p(keys), p(un), p(ur), p(dp), p(d), n, p(bw), p(bv), len(bw), p(fn), fo.ctypes.data if len(fn) else None,
                                      p(fi), len(fn), out.ctypes.data, cap, C.byref(w)))
    return out[:w.value].tobytes()


def decode(blob):
    """-> dict(keys, keys_un, u_right, depth, desc, bow=(words, vals), fv=(nodes, off, idx), same_un)"""
    buf = np.frombuffer(blob, np.uint8)
    v = [C.c_int() for _ in range(5)]
    check(_L().corb_kf_payload_info(buf.ctypes.data, len(buf), *[C.byref(x) for x in v]))
    n, nb, nf, ni, same = [x.value for x in v]
    keys, un = np.zeros(n, KP_DTYPE), np.zeros(n, KP_DTYPE)
    ur, dp, d = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 32), np.uint8)
    bw, bv = np.zeros(nb, np.uint32), np.zeros(nb, np.float64)
    fn, fo, fi = np.zeros(nf, np.uint32), np.zeros(nf + 1, np.int32), np.zeros(ni, np.uint32)
    check(_L().corb_kf_payload_decode(buf.ctypes.data, len(buf), *[a.ctypes.data for a in (keys, un, ur, dp, d, bw, bv, fn, fo, fi)]))
    return {"keys": keys, "keys_un": un, "u_right": ur, "depth": dp, "desc": d, "bow": (bw, bv), "fv": (fn, fo, fi), "same_un": bool(same)}
