"""Host-side mirror of the ORBmatcher entry points on the hot path (reference: corbslam_client/include/ORBmatcher.h:41-67,
src/ORBmatcher.cc:162-423,657-790,1792-1808) over the C ABI. The reference walks KeyFrame/Frame objects; here a
`BowFeatures` bundle carries exactly the members those functions read (mDescriptors, mFeatVec, MapPoint liveness,
keypoint angles), which is also what the C++ shim flattens."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import BowSide, check, lib

KF_FRAME, KF_SERVER, KF_KF = 0, 1, 2


class BowFeatures:
    """What SearchByBoW reads from one KeyFrame / Frame.

    desc (n,32) u8 = mDescriptors; fv = (nodes, off, idx) = mFeatVec flattened; valid[n] = MapPoint* != NULL and
    !isBad() (GetMapPointMatches); angles[n] = mvKeysUn[i].angle or mvKeys[i].angle as the variant prescribes."""

    def __init__(self, desc, fv_nodes, fv_off, fv_idx, valid=None, angles=None):
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.n = len(self.desc)
        self.fv_nodes = np.ascontiguousarray(fv_nodes, np.uint32)
        self.fv_off = np.ascontiguousarray(fv_off, np.int32)
        self.fv_idx = np.ascontiguousarray(fv_idx, np.uint32)
        self.valid = None if valid is None else np.ascontiguousarray(valid, np.uint8)
        self.angles = None if angles is None else np.ascontiguousarray(angles, np.float32)

    def c_side(self):
        s = BowSide()
        s.desc = self.desc.ctypes.data
        s.n = self.n
        s.fv_nodes = self.fv_nodes.ctypes.data
        s.fv_off = self.fv_off.ctypes.data
        s.fv_idx = self.fv_idx.ctypes.data
        s.fv_n = len(self.fv_nodes)
        s.valid = self.valid.ctypes.data if self.valid is not None else None
        s.angles = self.angles.ctypes.data if self.angles is not None else None
        return s


class ORBmatcher:
    TH_LOW = 50
    TH_HIGH = 100
    HISTO_LENGTH = 30

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        h = C.c_void_p()
        check(lib().corb_matcher_create(int(device), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().corb_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    # ---- SearchByProjection (ORBmatcher.h:51,55; ORBmatcher.cc:44-131, 1470-1614)
    def SearchByProjectionLastFrame(self, cur, last_valid, last_xyz, last_mp_desc, last_octave, last_angle, Tlw, th, bMono=False,
                                    last_blocks=None):
        """SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono). `cur` is a frame.FrameView; the
        LastFrame members are passed flattened (see include/corb_b200.h). Returns (match[cur.n] -> last index or -1, nmatches)."""
        u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
        valid, blocks = u8(last_valid), u8(last_blocks)
        xyz = np.ascontiguousarray(last_xyz, np.float32)
        desc = np.ascontiguousarray(last_mp_desc, np.uint8)
        octv = np.ascontiguousarray(last_octave, np.int32)
        ang = np.ascontiguousarray(last_angle, np.float32)
        T = np.ascontiguousarray(np.asarray(Tlw, np.float32).reshape(-1)[:12])
        match = np.empty(cur.n, np.int32)
        nm = C.c_int32()
        cs = cur.c_struct()
        check(lib().corb_search_by_projection_last(self._h, C.byref(cs), len(valid), valid.ctypes.data,
                                                   blocks.ctypes.data if blocks is not None else None, xyz.ctypes.data,
                                                   desc.ctypes.data, octv.ctypes.data, ang.ctypes.data, T.ctypes.data, float(th),
                                                   int(bool(bMono)), int(self.mbCheckOrientation), match.ctypes.data, C.byref(nm)))
        return match, nm.value

    def SearchByProjectionMapPoints(self, F, in_view, proj, level, view_cos, mp_desc, th=1.0, blocks=None):
        """SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th): proj[:, 0:3] = mTrackProjX, mTrackProjY,
        mTrackProjXR; level = mnTrackScaleLevel; view_cos = mTrackViewCos. Returns (match[F.n] -> map point index or -1, n)."""
        u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
        iv, bl = u8(in_view), u8(blocks)
        pr = np.ascontiguousarray(proj, np.float32)
        lv = np.ascontiguousarray(level, np.int32)
        vc = np.ascontiguousarray(view_cos, np.float32)
        desc = np.ascontiguousarray(mp_desc, np.uint8)
        match = np.empty(F.n, np.int32)
        nm = C.c_int32()
        cs = F.c_struct()
        check(lib().corb_search_by_projection_map(self._h, C.byref(cs), len(iv), iv.ctypes.data, bl.ctypes.data if bl is not None else None,
                                                  pr.ctypes.data, lv.ctypes.data, vc.ctypes.data, desc.ctypes.data, float(th),
                                                  float(self.mfNNratio), match.ctypes.data, C.byref(nm)))
        return match, nm.value

    @staticmethod
    def DescriptorDistance(a, b):
        """Scalar form stays on the host like the reference's inline popcount (ORBmatcher.cc:1792-1808)."""
        x = np.bitwise_xor(np.ascontiguousarray(a, np.uint8).ravel(), np.ascontiguousarray(b, np.uint8).ravel())
        return int(np.unpackbits(x).sum())

    def DescriptorDistancePairs(self, A, B, pairs):
        A = np.ascontiguousarray(A, np.uint8).reshape(-1, 32)
        B = np.ascontiguousarray(B, np.uint8).reshape(-1, 32)
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        out = np.empty(len(pairs), np.int32)
        check(lib().corb_hamming_pairs(self._h, A.ctypes.data, len(A), B.ctypes.data, len(B), pairs.ctypes.data, len(pairs),
                                       out.ctypes.data))
        return out

    def _batch(self, variant, As, Bs):
        n = len(As)
        sa = (BowSide * n)(*[a.c_side() for a in As])
        sb = (BowSide * n)(*[b.c_side() for b in Bs])
        outs = [np.empty(max(1, a.n if variant == KF_KF else b.n), np.int32) for a, b in zip(As, Bs)]
        ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        nm = np.zeros(n, np.int32)
        check(lib().corb_bow_match_batch(self._h, variant, n, sa, sb, self.mfNNratio, int(self.mbCheckOrientation), ptrs,
                                         nm.ctypes.data_as(_lib.i32p)))
        return [(o[:(a.n if variant == KF_KF else b.n)], int(k)) for o, a, b, k in zip(outs, As, Bs, nm)]

    def SearchByBoWRecords(self, variant, recsA, validA, recsB, validB=None):
        """SearchByBoW between device-resident BowRecords (corb_bow_match_stores): validA / validB = lists of host byte masks
        (or None). -> [(match, nmatches)] like SearchByBoWBatch. The marshalling is vectorised (one output block, pointer
        tables built with numpy): at ~70 us of kernel time per batch of 32 calls the Python side is what is left to trim."""
        n = len(recsA)
        keep = []

        def masks(valids):
            tab = np.zeros(n, np.uint64)
            if valids is not None:
                for i, v in enumerate(valids):
                    if v is not None:
                        v = np.ascontiguousarray(v, np.uint8)
                        keep.append(v)
                        tab[i] = v.__array_interface__["data"][0]
            keep.append(tab)
            return (C.c_void_p * n).from_buffer(tab)

        def handles(recs):
            tab = np.fromiter((r._h.value for r in recs), np.uint64, n)
            keep.append(tab)
            return (C.c_void_p * n).from_buffer(tab)

        sizes = np.fromiter(((a.n if variant == KF_KF else b.n) for a, b in zip(recsA, recsB)), np.int64, n)
        offs = np.zeros(n + 1, np.int64)
        np.cumsum(np.maximum(sizes, 1), out=offs[1:])
        out = np.empty(int(offs[-1]), np.int32)
        ptab = (out.__array_interface__["data"][0] + 4 * offs[:-1]).astype(np.uint64)
        nm = np.zeros(n, np.int32)
        check(lib().corb_bow_match_stores(self._h, variant, n, handles(recsA), masks(validA), handles(recsB), masks(validB),
                                          self.mfNNratio, int(self.mbCheckOrientation), (C.c_void_p * n).from_buffer(ptab),
                                          nm.ctypes.data_as(_lib.i32p)))
        return [(out[offs[i]:offs[i] + sizes[i]], int(nm[i])) for i in range(n)]

    def SearchByBoW(self, kf, frame):
        """SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches): -> (match[F.N] = KF feature index or -1, nmatches)."""
        return self._batch(KF_FRAME, [kf], [frame])[0]

    def SearchByBoWInServer(self, kf, f):
        return self._batch(KF_SERVER, [kf], [f])[0]

    def SearchByBoWKF(self, kf1, kf2):
        """SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12): -> (match[N1] = KF2 feature index or -1, nmatches)."""
        return self._batch(KF_KF, [kf1], [kf2])[0]

    def SearchByBoWBatch(self, variant, kfs, frames):
        """ncalls independent calls in one launch (relocalisation / map-fusion candidates)."""
        return self._batch(variant, kfs, frames)
