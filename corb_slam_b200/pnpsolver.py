"""Host-side mirror of PnPsolver (reference: corbslam_client/include/PnPsolver.h:64-78, src/PnPsolver.cc:66-383) over the
C ABI (corb_pnp_ransac_params, corb_pnp_iterate_batch). Same names, argument meaning and results as the reference class;
`PnPsolver.iterate_batch` is the form the relocalisation / map-fusion loops call (Tracking.cc:1432-1440,
MapFusion.cpp:720-728 iterate every candidate's solver in turn - here all candidates go to the GPU in one call).

The reference draws its minimal sets with DUtils::Random::RandomInt, i.e. from the process-global rand() stream
(Random.cpp:47-50). The mirror does the same by default (glibc rand() through ctypes) and hands the draws to the
library, which makes the GPU result a pure function of its arguments; pass `rand=` to supply another source.

Deviation in how much of the rand() stream is consumed: the first iterate() of a solver draws 4 values for EVERY iteration
up to max(mRansacMaxIts, done + nIterations), because the whole batch is evaluated at once; the reference draws 4 per
iteration it actually executes and stops at the first successful Refine(). A single solver sees exactly the reference's
draws (tests/test_ref_gpu.py); with several candidates in one process (Tracking.cc:1432, MapFusion.cpp:720) every solver
after the first starts further down the stream than it would in the reference - still a valid RANSAC, not the same one."""
import ctypes as C
import threading

import numpy as np

from ._lib import PnpProblem, PnpResult, check, lib

RAND_MAX = 2147483647
_libc = None


def _glibc_rand():
    global _libc
    if _libc is None:
        _libc = C.CDLL(None)
        _libc.rand.restype = C.c_int
    return _libc.rand()


def random_int(lo, hi, rand=_glibc_rand):
    """DUtils::Random::RandomInt (Random.cpp:47-50)."""
    d = hi - lo + 1
    return int((float(rand()) / (float(RAND_MAX) + 1.0)) * d) + lo


_tls = threading.local()


def _matcher(device):
    """One corb_matcher handle per (thread, device): a handle owns a stream and the staging arena of corb_pnp_iterate_batch and
    must not be used from two threads at once (include/corb_b200.h), while the reference runs PnPsolver::iterate concurrently
    from Tracking::Relocalization and the server's MapFusion thread; ctypes releases the GIL during the call."""
    handles = getattr(_tls, "handles", None)
    if handles is None:
        handles = _tls.handles = {}
    h = handles.get(device)
    if h is None:
        h = C.c_void_p()
        check(lib().corb_matcher_create(int(device), C.byref(h)))
        handles[device] = h
    return h


class PnPsolver:
    def __init__(self, keys_un, octaves, level_sigma2, K, map_point_xyz, map_point_valid, device=0, rand=_glibc_rand):
        """PnPsolver(const Frame &F, const vector<MapPoint*> &vpMapPointMatches) (:66-111; the KeyFrame overload :113-151 is
        identical): keys_un (M,2) = F.mvKeysUn[i].pt, octaves (M) = .octave, level_sigma2 = F.mvLevelSigma2,
        K = (F.fx, F.fy, F.cx, F.cy), map_point_xyz (M,3) = GetWorldPos() of vpMapPointMatches[i],
        map_point_valid (M) = pMP && !pMP->isBad()."""
        valid = np.asarray(map_point_valid).astype(bool)
        self.n_matches = len(valid)
        self.mvKeyPointIndices = np.nonzero(valid)[0].astype(np.int32)
        self.mvP2D = np.ascontiguousarray(np.asarray(keys_un, np.float32).reshape(-1, 2)[valid])
        self.mvP3Dw = np.ascontiguousarray(np.asarray(map_point_xyz, np.float32).reshape(-1, 3)[valid])
        self.mvSigma2 = np.ascontiguousarray(np.asarray(level_sigma2, np.float32)[np.asarray(octaves)[valid]])
        self.fu, self.fv, self.uc, self.vc = [np.float32(v) for v in K]
        self.N = len(self.mvP2D)
        self.device = int(device)
        self._rand = rand
        self._draws = np.zeros((0, 4), np.int32)
        self.mnIterations = 0
        self.SetRansacParameters()

    def SetRansacParameters(self, probability=0.99, minInliers=8, maxIterations=300, minSet=4, epsilon=0.4, th2=5.991):
        """:163-198. The adjustment of mRansacMinInliers / mRansacMaxIts is the library's (host scalar logic)."""
        a, b = C.c_int(0), C.c_int(0)
        check(lib().corb_pnp_ransac_params(self.N, float(probability), int(minInliers), int(maxIterations), int(minSet), float(epsilon),
                                           C.byref(a), C.byref(b)))
        self.mRansacMinInliers, self.mRansacMaxIts, self.mRansacMinSet = a.value, b.value, int(minSet)
        self.mvMaxError = np.ascontiguousarray(self.mvSigma2 * np.float32(th2), np.float32)

    def _need_draws(self, it_end):
        """RandomInt(0, vAvailableIndices.size()-1) for the four picks of every iteration up to it_end (:231-233).

        Deviation from the reference, by design: the draws for ALL iterations up to it_end are taken from the process-global
        rand() stream up front (the batched kernel evaluates every hypothesis), while the reference consumes four per EXECUTED
        iteration and stops drawing at an early Refine() return. A single solver that runs to the end sees the reference's
        draws; with several candidate solvers in one process (Tracking.cc:1432, MapFusion.cpp:720) the solvers after the first
        start at a later stream position than they would in the reference process. The result is still a valid RANSAC over
        the same distribution; bit parity with the reference holds for given draws (set_draws, and the tests that replay the
        reference's own rand() stream through one solver)."""
        have = len(self._draws)
        if it_end > have:
            new = np.empty((it_end - have, 4), np.int32)
            for i in range(len(new)):
                for k in range(4):
                    new[i, k] = random_int(0, self.N - k - 1, self._rand)
            self._draws = np.ascontiguousarray(np.concatenate([self._draws, new]))

    def set_draws(self, draws):
        """Replace the random source by explicit RandomInt values, draws[it, k] (tests, replays)."""
        self._draws = np.ascontiguousarray(draws, np.int32).reshape(-1, 4)

    def _problem(self, nIterations):
        p = PnpProblem()
        p.n = self.N
        p.p2d, p.p3d, p.max_err = self.mvP2D.ctypes.data, self.mvP3Dw.ctypes.data, self.mvMaxError.ctypes.data
        p.fx, p.fy, p.cx, p.cy = float(self.fu), float(self.fv), float(self.uc), float(self.vc)
        p.min_inliers, p.max_its = self.mRansacMinInliers, self.mRansacMaxIts
        p.iterations_done, p.n_iterations = self.mnIterations, int(nIterations)
        if self.N >= self.mRansacMinInliers:
            self._need_draws(max(self.mRansacMaxIts, self.mnIterations + int(nIterations)))
        p.draws = self._draws.ctypes.data
        return p

    def _unpack(self, r, inl):
        self.mnIterations = r.iterations
        vbInliers = np.zeros(self.n_matches, bool)
        if r.status:
            vbInliers[self.mvKeyPointIndices[inl[:self.N].astype(bool)]] = True
            Tcw = np.array(r.Tcw, np.float32).reshape(4, 4)
        else:
            Tcw = None
        return Tcw, bool(r.no_more), vbInliers, int(r.n_inliers)

    @staticmethod
    def pack_batch(solvers, nIterations):
        """The C arguments of one corb_pnp_iterate_batch call (what the C++ shim's Fill() builds)."""
        n = len(solvers)
        probs = (PnpProblem * n)(*[s._problem(nIterations) for s in solvers])
        res = (PnpResult * n)()
        bufs = [np.zeros(max(s.N, 1), np.uint8) for s in solvers]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        return n, probs, res, ptrs, bufs

    @staticmethod
    def call_batch(device, packed):
        n, probs, res, ptrs, _ = packed
        check(lib().corb_pnp_iterate_batch(_matcher(device), n, probs, res, ptrs))

    @staticmethod
    def iterate_batch(solvers, nIterations):
        """iterate(nIterations, ...) of every solver, one GPU call. Returns a list of (Tcw | None, bNoMore, vbInliers, nInliers)."""
        if not solvers:
            return []
        packed = PnPsolver.pack_batch(solvers, nIterations)
        PnPsolver.call_batch(solvers[0].device, packed)
        return [s._unpack(packed[2][i], packed[4][i]) for i, s in enumerate(solvers)]

    def iterate(self, nIterations):
        """cv::Mat iterate(int nIterations, bool &bNoMore, vector<bool> &vbInliers, int &nInliers) (:206-300)."""
        return PnPsolver.iterate_batch([self], nIterations)[0]

    def find(self):
        """cv::Mat find(vector<bool> &vbInliers, int &nInliers) (:200-204)."""
        Tcw, _, vbInliers, nInliers = self.iterate(self.mRansacMaxIts)
        return Tcw, vbInliers, nInliers
