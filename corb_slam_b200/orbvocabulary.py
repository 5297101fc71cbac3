"""Host-side mirror of ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> for the calls on the hot path
(reference: corbslam_client/include/ORBVocabulary.h:31-32; Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1259,
1338-1424; ScoringObject.cpp:23-68) over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


class ORBVocabulary:
    def __init__(self, device=0):
        self._h = None
        self.device = int(device)

    def loadFromTextFile(self, filename):
        h = C.c_void_p()
        try:
            check(lib().corb_voc_load_text(str(filename).encode(), self.device, C.byref(h)))
        except _lib.CorbError as e:
            if e.status == _lib.ERR_IO:
                return False  # the reference returns false on a malformed file (:1361-1365)
            raise
        self._set(h)
        return True

    @classmethod
    def from_arrays(cls, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0, device=0):
        self = cls(device)
        parent = np.ascontiguousarray(parent, np.int32)
        is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8)
        weight = np.ascontiguousarray(weight, np.float64)
        h = C.c_void_p()
        check(lib().corb_voc_create(k, L, scoring, weighting, len(parent), parent.ctypes.data, is_leaf.ctypes.data,
                                    desc.ctypes.data, weight.ctypes.data, self.device, C.byref(h)))
        self._set(h)
        return self

    def _set(self, h):
        self.close()
        self._h = h
        v = [C.c_int32() for _ in range(6)]
        check(lib().corb_voc_info(h, *[C.byref(x) for x in v]))
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = [x.value for x in v]

    def close(self):
        if getattr(self, "_h", None):
            lib().corb_voc_destroy(self._h)
            self._h = None

    __del__ = close

    def empty(self):
        return self._h is None

    def size(self):
        return self.n_words

    def transform_features(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        w = np.empty(n, np.uint32); wt = np.empty(n, np.float64); nid = np.empty(n, np.uint32)
        check(lib().corb_voc_transform_features(self._h, desc.ctypes.data, n, levelsup, w.ctypes.data, wt.ctypes.data,
                                                nid.ctypes.data))
        return w, wt, nid

    def transform(self, desc, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup) -> (bow_words, bow_vals, fv_nodes, fv_off, fv_idx)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        bw = np.empty(max(n, 1), np.uint32); bv = np.empty(max(n, 1), np.float64)
        fn = np.empty(max(n, 1), np.uint32); fo = np.empty(n + 1, np.int32); fi = np.empty(max(n, 1), np.uint32)
        nb, nf = C.c_int32(), C.c_int32()
        check(lib().corb_voc_transform(self._h, desc.ctypes.data, n, levelsup, bw.ctypes.data, bv.ctypes.data, C.byref(nb),
                                       fn.ctypes.data, fo.ctypes.data, fi.ctypes.data, C.byref(nf)))
        g = nf.value
        return bw[:nb.value].copy(), bv[:nb.value].copy(), fn[:g].copy(), fo[:g + 1].copy(), fi[:fo[g]].copy()

    def score(self, v1, v2):
        """score(BowVector v1, BowVector v2); v = (words, vals)."""
        return float(self.score_batch(v1, [v2])[0])

    def score_batch(self, query, candidates):
        qw = np.ascontiguousarray(query[0], np.uint32); qv = np.ascontiguousarray(query[1], np.float64)
        cws = [np.ascontiguousarray(c[0], np.uint32) for c in candidates]
        cvs = [np.ascontiguousarray(c[1], np.float64) for c in candidates]
        n = len(candidates)
        pw = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in cws])
        pv = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in cvs])
        cn = np.array([len(a) for a in cws], np.int32)
        out = np.empty(n, np.float64)
        check(lib().corb_bow_score_batch(self._h, qw.ctypes.data, qv.ctypes.data, len(qw), n, pw, pv, cn.ctypes.data,
                                         out.ctypes.data))
        return out


class BowRecord:
    """Device-resident matching record of a Frame / KeyFrame (corb_bow_store): descriptors, angles, BowVector and
    FeatureVector stay in HBM (reference: Frame::ComputeBoW, Frame.cc:399-406; KeyFrame::ComputeBoW, KeyFrame.cc:70-77)."""

    def __init__(self, capacity=2048, device=0):
        self._h = C.c_void_p()
        check(lib().corb_bow_store_create(int(device), int(capacity), C.byref(self._h)))
        self.device, self.capacity, self.n = int(device), int(capacity), 0

    def close(self):
        if getattr(self, "_h", None):
            lib().corb_bow_store_destroy(self._h)
            self._h = None

    __del__ = close

    def from_extractor(self, extractor, voc, levelsup=4):
        """corb_frame_bow: record <- the extractor's last (device-resident) result."""
        check(lib().corb_frame_bow(extractor._h, voc._h, int(levelsup), self._h))
        self.n = int(lib().corb_bow_store_features(self._h))  # known when the fill is enqueued: no wait here
        return self

    def from_device(self, voc, d_desc, n, levelsup=4, d_kps=None, stream=None):
        check(lib().corb_bow_store_fill(self._h, voc._h, d_kps, d_desc, int(n), int(levelsup), stream))
        self.n = int(n)
        return self

    def sync(self):
        """corb_bow_store_sync: wait for the record's fill (from_extractor / from_device only enqueue it)."""
        check(lib().corb_bow_store_sync(self._h))
        return self

    def side(self, d_valid=None):
        """corb_bow_side of device pointers (for ORBmatcher.SearchByBoWDevice) and the BowVector length."""
        s, nb = _lib.BowSide(), C.c_int32()
        check(lib().corb_bow_store_side(self._h, d_valid, C.byref(s), C.byref(nb)))
        return s, nb.value

    def download(self):
        """-> (bow_words, bow_vals, fv_nodes, fv_off, fv_idx) like ORBVocabulary.transform."""
        cap = self.capacity
        bw, bv = np.empty(cap, np.uint32), np.empty(cap, np.float64)
        fn, fo, fi = np.empty(cap, np.uint32), np.empty(cap + 1, np.int32), np.empty(cap, np.uint32)
        nb, nf = C.c_int32(), C.c_int32()
        check(lib().corb_bow_store_download(self._h, bw.ctypes.data, bv.ctypes.data, C.byref(nb), fn.ctypes.data, fo.ctypes.data,
                                            fi.ctypes.data, C.byref(nf)))
        g = nf.value
        return bw[:nb.value].copy(), bv[:nb.value].copy(), fn[:g].copy(), fo[:g + 1].copy(), fi[:fo[g] if g else 0].copy()

    @staticmethod
    def score(voc, query, candidates):
        """voc.score(query, c) for every candidate record, one launch."""
        n = len(candidates)
        arr = (C.c_void_p * max(n, 1))(*[c._h for c in candidates])
        out = np.empty(n, np.float64)
        check(lib().corb_bow_score_stores(voc._h, query._h, n, arr, out.ctypes.data))
        return out
