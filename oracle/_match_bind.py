"""ctypes signatures + numpy wrappers of oracle/match_oracle.cpp (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C

import numpy as np

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_L = None


def bind(L):
    global _L
    _L = L
    L.oracle_hamming256.argtypes = [_u8p, _u8p]
    L.oracle_search_by_bow.argtypes = [C.c_int, _u8p, C.c_int, _u8p, C.c_int, _u32p, _i32p, _u32p, C.c_int, _u32p, _i32p, _u32p,
                                       C.c_int, _u8p, _u8p, _f32p, _f32p, C.c_float, C.c_int, _i32p]
    L.oracle_voc_create.restype = C.c_void_p
    L.oracle_voc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _u8p, _u8p, _f64p]
    L.oracle_voc_load_text.restype = C.c_void_p
    L.oracle_voc_load_text.argtypes = [C.c_char_p]
    L.oracle_voc_destroy.argtypes = [C.c_void_p]
    L.oracle_voc_info.argtypes = [C.c_void_p] + [_i32p] * 6
    L.oracle_voc_export.argtypes = [C.c_void_p, _i32p, _u8p, _u8p, _f64p]
    L.oracle_voc_transform.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _u32p, _f64p, _u32p]
    L.oracle_bow_build.argtypes = [C.c_int, _u32p, _f64p, _u32p, _u32p, _f64p, _u32p, _i32p, _u32p, _i32p]
    L.oracle_bow_score_l1.restype = C.c_double
    L.oracle_bow_score_l1.argtypes = [C.c_int, _u32p, _f64p, C.c_int, _u32p, _f64p]


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _lib():
    from . import lib
    lib()
    return _L


def hamming256(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return _lib().oracle_hamming256(_p(a, _u8p), _p(b, _u8p))


class Side:
    """One side of SearchByBoW: descriptors, FeatureVector (CSR), MapPoint liveness, keypoint angles."""

    def __init__(self, desc, fv_nodes, fv_off, fv_idx, valid=None, angles=None):
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.fv_nodes = np.ascontiguousarray(fv_nodes, np.uint32)
        self.fv_off = np.ascontiguousarray(fv_off, np.int32)
        self.fv_idx = np.ascontiguousarray(fv_idx, np.uint32)
        n = len(self.desc)
        self.valid = np.ones(n, np.uint8) if valid is None else np.ascontiguousarray(valid, np.uint8)
        self.angles = np.zeros(n, np.float32) if angles is None else np.ascontiguousarray(angles, np.float32)
        self.n = n


def search_by_bow(variant, A, B, nnratio, check_ori):
    n_out = A.n if variant == 2 else B.n
    match = np.empty(max(n_out, 1), np.int32)
    nm = _lib().oracle_search_by_bow(variant, _p(A.desc, _u8p), A.n, _p(B.desc, _u8p), B.n, _p(A.fv_nodes, _u32p),
                                     _p(A.fv_off, _i32p), _p(A.fv_idx, _u32p), len(A.fv_nodes), _p(B.fv_nodes, _u32p),
                                     _p(B.fv_off, _i32p), _p(B.fv_idx, _u32p), len(B.fv_nodes), _p(A.valid, _u8p),
                                     _p(B.valid, _u8p), _p(A.angles, _f32p), _p(B.angles, _f32p), float(nnratio),
                                     int(bool(check_ori)), _p(match, _i32p))
    return match[:n_out], nm


class Vocabulary:
    def __init__(self, handle):
        if not handle:
            raise ValueError("vocabulary could not be created/loaded")
        self._h = handle
        v = [C.c_int32() for _ in range(6)]
        _lib().oracle_voc_info(self._h, *[C.byref(x) for x in v])
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = [x.value for x in v]

    @classmethod
    def from_arrays(cls, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0):
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8); weight = np.ascontiguousarray(weight, np.float64)
        return cls(_lib().oracle_voc_create(k, L, scoring, weighting, len(parent), _p(parent, _i32p), _p(is_leaf, _u8p),
                                            _p(desc, _u8p), _p(weight, _f64p)))

    @classmethod
    def load_text(cls, path):
        return cls(_lib().oracle_voc_load_text(path.encode()))

    def __del__(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.oracle_voc_destroy(self._h)
            self._h = None

    def export(self):
        n = self.n_nodes - 1
        parent = np.empty(n, np.int32); leaf = np.empty(n, np.uint8); desc = np.empty((n, 32), np.uint8)
        weight = np.empty(n, np.float64)
        _lib().oracle_voc_export(self._h, _p(parent, _i32p), _p(leaf, _u8p), _p(desc, _u8p), _p(weight, _f64p))
        return parent, leaf, desc, weight

    def transform_features(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        w = np.empty(n, np.uint32); wt = np.empty(n, np.float64); nid = np.empty(n, np.uint32)
        rc = _lib().oracle_voc_transform(self._h, _p(desc, _u8p), n, levelsup, _p(w, _u32p), _p(wt, _f64p), _p(nid, _u32p))
        assert rc == 0
        return w, wt, nid

    def transform(self, desc, levelsup=4):
        """-> (bow_words, bow_vals, fv_nodes, fv_off, fv_idx), i.e. BowVector and FeatureVector flattened."""
        w, wt, nid = self.transform_features(desc, levelsup)
        return bow_build(w, wt, nid)

    @staticmethod
    def score(b1, b2):
        return bow_score_l1(b1[0], b1[1], b2[0], b2[1])


def bow_build(word_id, weight, node_id):
    n = len(word_id)
    ow = np.empty(max(n, 1), np.uint32); ov = np.empty(max(n, 1), np.float64)
    fn = np.empty(max(n, 1), np.uint32); fo = np.empty(n + 1, np.int32); fi = np.empty(max(n, 1), np.uint32)
    nfv = C.c_int32()
    word_id = np.ascontiguousarray(word_id, np.uint32); weight = np.ascontiguousarray(weight, np.float64)
    node_id = np.ascontiguousarray(node_id, np.uint32)
    m = _lib().oracle_bow_build(n, _p(word_id, _u32p), _p(weight, _f64p), _p(node_id, _u32p), _p(ow, _u32p), _p(ov, _f64p),
                                _p(fn, _u32p), _p(fo, _i32p), _p(fi, _u32p), C.byref(nfv))
    g = nfv.value
    return ow[:m].copy(), ov[:m].copy(), fn[:g].copy(), fo[:g + 1].copy(), fi[:fo[g]].copy()


def bow_score_l1(w1, v1, w2, v2):
    w1 = np.ascontiguousarray(w1, np.uint32); v1 = np.ascontiguousarray(v1, np.float64)
    w2 = np.ascontiguousarray(w2, np.uint32); v2 = np.ascontiguousarray(v2, np.float64)
    return _lib().oracle_bow_score_l1(len(w1), _p(w1, _u32p), _p(v1, _f64p), len(w2), _p(w2, _u32p), _p(v2, _f64p))


def random_vocabulary(k=10, L=3, seed=0, stop_fraction=0.02):
    """Synthetic vocabulary tree (k-ary, L levels) in DBoW2 file order: random 256-bit node descriptors, idf-like leaf
    weights, a few stopped (weight 0) words. Returns (k, L, parent, is_leaf, desc, weight)."""
    rng = np.random.default_rng(seed)
    parent, leaf = [], []
    frontier = [0]
    nid = 0
    for level in range(1, L + 1):
        nxt = []
        for p in frontier:
            for _ in range(k):
                nid += 1
                parent.append(p)
                leaf.append(1 if level == L else 0)
                nxt.append(nid)
        frontier = nxt
    n = nid
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    weight = np.where(np.array(leaf) > 0, rng.uniform(0.5, 9.0, n), 0.0)
    weight[(rng.random(n) < stop_fraction) & (np.array(leaf) > 0)] = 0.0
    return k, L, np.array(parent, np.int32), np.array(leaf, np.uint8), desc, weight
