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
        vp = C.c_void_p
        L.ref_set_monotone_nodes.argtypes = [C.c_int]
        L.ref_frame_stereo.restype = vp
        L.ref_frame_stereo.argtypes = [vp, vp, _u8p, _u8p, C.c_int, C.c_int, C.c_int] + [C.c_float] * 6
        L.ref_frame_destroy.argtypes = [vp]
        L.ref_frame_counts.argtypes = [vp, _i32p, _i32p]
        L.ref_frame_results.argtypes = [vp, vp, _u8p, vp, _u8p, _f32p, _f32p]
        L.ref_frame_grid.argtypes = [vp, _i32p, _i32p, _f32p]
        L.ref_stereo_matches.argtypes = [vp, vp, vp, _u8p, C.c_int, vp, _u8p, C.c_int, C.c_float, C.c_float, _f32p, _f32p]
        L.ref_descriptor_distance.argtypes = [_u8p, _u8p]
        L.ref_search_by_bow.argtypes = [C.c_int, _u8p, C.c_int, _u8p, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp,
                                        C.c_float, C.c_int, vp]
        L.ref_search_by_projection_last.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp]
        L.ref_search_by_projection_map.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp]
        L.ref_features_in_area.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, vp]
        L.ref_voc_load_text.restype = vp
        L.ref_voc_load_text.argtypes = [C.c_char_p]
        L.ref_voc_destroy.argtypes = [vp]
        L.ref_voc_info.argtypes = [vp, _i32p, _i32p, _i32p]
        L.ref_voc_transform.argtypes = [vp, _u8p, C.c_int, C.c_int, vp, vp, vp, vp, vp, _i32p]
        L.ref_voc_score.restype = C.c_double
        L.ref_voc_score.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp]
        L.ref_srand.argtypes = [C.c_uint]
        L.ref_rand_draws.argtypes = [C.c_uint, C.c_int, C.c_int, vp]
        L.ref_pnp_create.restype = vp
        L.ref_pnp_create.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int] + [C.c_float] * 4
        L.ref_pnp_destroy.argtypes = [vp]
        L.ref_pnp_set_ransac.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _i32p, _i32p]
        L.ref_pnp_iterations.argtypes = [vp]
        L.ref_pnp_iterate.argtypes = [vp, C.c_int, _i32p, vp, _i32p, vp]
        L.ref_cv_gemm_f32.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_double, vp, C.c_int, C.c_int, C.c_double, C.c_int, vp]
        L.ref_expr_minus_Rt_t.argtypes = [vp, vp, vp]
        L.ref_expr_Rx_plus_t.argtypes = [vp, vp, vp, vp]
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


def vocabulary_text(stripped=False):
    """Path of the reference's ORBvoc.txt (corbslam_client/Vocabulary/ORBvoc.txt.tar.gz, k = 10, L = 6, 1 082 073 nodes),
    untarred once into a temp directory. `stripped`: without the trailing newline - the reference's loadFromTextFile
    (TemplatedVocabulary.h:1378-1395) reads one more line after the last node and then uses `pid` / `nIsLeaf` that the failed
    extraction never wrote (indeterminate values, a crash under -O3 here), so the reference loader gets the same nodes
    without the blank tail; the oracle and corb_voc_load_text ignore the blank line by definition."""
    import tarfile
    import tempfile
    cache = os.path.join(tempfile.gettempdir(), "corb_voc_%d" % os.getuid())
    full = os.path.join(cache, "ORBvoc.txt")
    if not os.path.exists(full):
        src = [p for p in (os.path.join(_HERE, "_ref", "ORBvoc.txt.tar.gz"), os.path.join(REFERENCE, "Vocabulary", "ORBvoc.txt.tar.gz"))
               if os.path.exists(p)]
        if not src:
            raise FileNotFoundError("ORBvoc.txt.tar.gz (run `make -C oracle/refbuild` where /root/reference is mounted)")
        os.makedirs(cache, exist_ok=True)
        with tarfile.open(src[0]) as t:
            t.extract("ORBvoc.txt", cache + ".tmp", filter="data")
        os.replace(os.path.join(cache + ".tmp", "ORBvoc.txt"), full)
    if not stripped:
        return full
    cut = os.path.join(cache, "ORBvoc_stripped.txt")
    if not os.path.exists(cut):
        data = open(full, "rb").read().rstrip()
        with open(cut + ".tmp", "wb") as f:
            f.write(data)
        os.replace(cut + ".tmp", cut)
    return cut


def set_monotone_nodes(on):
    """Creation-ordered std::list<ExtractorNode> node addresses (default) vs plain malloc (oracle/refbuild/ref_alloc.cpp)."""
    lib().ref_set_monotone_nodes(int(bool(on)))


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Frame:
    """ORB_SLAM2::Frame built by its stereo constructor (Frame.cc:60-124): ExtractORB on two threads,
    ComputeStereoMatches, image bounds, AssignFeaturesToGrid - all the reference's own code."""

    def __init__(self, ex_left, ex_right, left, right, fx, fy, cx, cy, bf, th_depth=35.0):
        left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
        h, w = left.shape
        self._keep = (ex_left, ex_right)
        self._h = lib().ref_frame_stereo(ex_left._h, ex_right._h, _p(left, _u8p), _p(right, _u8p), w, h, left.strides[0],
                                         fx, fy, cx, cy, bf, th_depth)
        nl, nr = C.c_int32(), C.c_int32()
        lib().ref_frame_counts(self._h, C.byref(nl), C.byref(nr))
        nl, nr = nl.value, nr.value
        self.keys = np.zeros(nl, KP_DTYPE); self.keys_right = np.zeros(nr, KP_DTYPE)
        self.desc = np.zeros((nl, 32), np.uint8); self.desc_right = np.zeros((nr, 32), np.uint8)
        self.u_right = np.zeros(nl, np.float32); self.depth = np.zeros(nl, np.float32)
        lib().ref_frame_results(self._h, _ptr(self.keys), _p(self.desc, _u8p), _ptr(self.keys_right), _p(self.desc_right, _u8p),
                                _p(self.u_right, _f32p), _p(self.depth, _f32p))
        self.grid_off = np.zeros(64 * 48 + 1, np.int32); self.grid_idx = np.zeros(max(nl, 1), np.int32)
        self.bounds = np.zeros(6, np.float32)
        lib().ref_frame_grid(self._h, _p(self.grid_off, _i32p), _p(self.grid_idx, _i32p), _p(self.bounds, _f32p))
        self.grid_idx = self.grid_idx[:self.grid_off[-1]]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_frame_destroy(self._h)
            self._h = None


def stereo_matches(ex_left, ex_right, kps_l, desc_l, kps_r, desc_r, mbf, mb):
    """Frame::ComputeStereoMatches (Frame.cc:470-644) on the extractors' last pyramids -> (mvuRight, mvDepth)."""
    kl = np.ascontiguousarray(kps_l); kr = np.ascontiguousarray(kps_r)
    dl = np.ascontiguousarray(desc_l, np.uint8); dr = np.ascontiguousarray(desc_r, np.uint8)
    ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32)
    lib().ref_stereo_matches(ex_left._h, ex_right._h, _ptr(kl), _p(dl, _u8p), len(kl), _ptr(kr), _p(dr, _u8p), len(kr), float(mbf),
                             float(mb), _p(ur, _f32p), _p(dp, _f32p))
    return ur, dp


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().ref_descriptor_distance(_p(a, _u8p), _p(b, _u8p))


def search_by_bow(variant, A, B, nnratio, check_ori):
    """ORBmatcher::SearchByBoW(KF, Frame) / SearchByBoWInServer / SearchByBoW(KF, KF) on oracle._match_bind.Side objects."""
    n_out = A.n if variant == 2 else B.n
    match = np.full(max(n_out, 1), -1, np.int32)
    nm = lib().ref_search_by_bow(variant, _p(A.desc, _u8p), A.n, _p(B.desc, _u8p), B.n, _ptr(A.fv_nodes), _ptr(A.fv_off), _ptr(A.fv_idx),
                                 len(A.fv_nodes), _ptr(B.fv_nodes), _ptr(B.fv_off), _ptr(B.fv_idx), len(B.fv_nodes), _ptr(A.valid),
                                 _ptr(B.valid), _ptr(A.angles), _ptr(B.angles), float(nnratio), int(bool(check_ori)), _ptr(match))
    return match[:n_out], nm


def search_by_projection_last(view_struct, n_cur, last_valid, last_blocks, last_xyz, last_mp_desc, last_octave, last_angle, Tlw, th,
                              mono, check_ori):
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    valid, blocks = u8(last_valid), u8(last_blocks)
    xyz = np.ascontiguousarray(last_xyz, np.float32); desc = np.ascontiguousarray(last_mp_desc, np.uint8)
    octv = np.ascontiguousarray(last_octave, np.int32); ang = np.ascontiguousarray(last_angle, np.float32)
    T = np.ascontiguousarray(np.asarray(Tlw, np.float32).reshape(-1)[:12])
    match = np.full(n_cur, -1, np.int32)
    n = lib().ref_search_by_projection_last(C.addressof(view_struct), len(valid), _ptr(valid), _ptr(blocks), _ptr(xyz), _ptr(desc),
                                            _ptr(octv), _ptr(ang), _ptr(T), float(th), int(mono), int(check_ori), _ptr(match))
    return match, n


def search_by_projection_map(view_struct, n_frame, in_view, blocks, proj, level, view_cos, mp_desc, th, nnratio):
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    iv, bl = u8(in_view), u8(blocks)
    pr = np.ascontiguousarray(proj, np.float32); lv = np.ascontiguousarray(level, np.int32)
    vc = np.ascontiguousarray(view_cos, np.float32); desc = np.ascontiguousarray(mp_desc, np.uint8)
    match = np.full(n_frame, -1, np.int32)
    n = lib().ref_search_by_projection_map(C.addressof(view_struct), len(iv), _ptr(iv), _ptr(bl), _ptr(pr), _ptr(lv), _ptr(vc),
                                           _ptr(desc), float(th), float(nnratio), _ptr(match))
    return match, n


class ORBVocabulary:
    """DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> itself (ORBVocabulary.h:31-32)."""

    def __init__(self, path):
        self._h = lib().ref_voc_load_text(os.fsencode(path))
        if not self._h:
            raise ValueError("loadFromTextFile failed: %s" % path)
        k, L, n = C.c_int32(), C.c_int32(), C.c_int32()
        lib().ref_voc_info(self._h, C.byref(k), C.byref(L), C.byref(n))
        self.k, self.L, self.n_words = k.value, L.value, n.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_voc_destroy(self._h)
            self._h = None

    def transform(self, desc, levelsup=4):
        """-> (bow_words, bow_vals, fv_nodes, fv_off, fv_idx): BowVector and FeatureVector in key order."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        ow = np.empty(max(n, 1), np.uint32); ov = np.empty(max(n, 1), np.float64)
        fn = np.empty(max(n, 1), np.uint32); fo = np.empty(n + 1, np.int32); fi = np.empty(max(n, 1), np.uint32)
        g = C.c_int32()
        m = lib().ref_voc_transform(self._h, _p(desc, _u8p), n, levelsup, _ptr(ow), _ptr(ov), _ptr(fn), _ptr(fo), _ptr(fi), C.byref(g))
        g = g.value
        return ow[:m].copy(), ov[:m].copy(), fn[:g].copy(), fo[:g + 1].copy(), fi[:fo[g]].copy()

    def score(self, b1, b2):
        w1 = np.ascontiguousarray(b1[0], np.uint32); v1 = np.ascontiguousarray(b1[1], np.float64)
        w2 = np.ascontiguousarray(b2[0], np.uint32); v2 = np.ascontiguousarray(b2[1], np.float64)
        return lib().ref_voc_score(self._h, len(w1), _ptr(w1), _ptr(v1), len(w2), _ptr(w2), _ptr(v2))


def rand_draws(seed, N, iters):
    """What DUtils::Random::RandomInt returns for the first `iters` RANSAC iterations after srand(seed)."""
    d = np.zeros((iters, 4), np.int32)
    lib().ref_rand_draws(int(seed), int(N), int(iters), _ptr(d))
    return d


class PnPsolver:
    """ORB_SLAM2::PnPsolver itself, fed through Frame / MapPoint objects (PnPsolver.cc:66-113)."""

    def __init__(self, p2d, octave, p3d, level_sigma2, fx, fy, cx, cy):
        self.p2d = np.ascontiguousarray(p2d, np.float32).reshape(-1, 2)
        self.octave = np.ascontiguousarray(octave, np.int32)
        self.p3d = np.ascontiguousarray(p3d, np.float32).reshape(-1, 3)
        s2 = np.ascontiguousarray(level_sigma2, np.float32)
        self.N = len(self.p2d)
        self._h = lib().ref_pnp_create(self.N, _ptr(self.p2d), _ptr(self.octave), _ptr(self.p3d), _ptr(s2), len(s2), fx, fy, cx, cy)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_pnp_destroy(self._h)
            self._h = None

    def SetRansacParameters(self, probability=0.99, minInliers=8, maxIterations=300, minSet=4, epsilon=0.4, th2=5.991):
        a, b = C.c_int32(), C.c_int32()
        lib().ref_pnp_set_ransac(self._h, probability, minInliers, maxIterations, minSet, epsilon, th2, C.byref(a), C.byref(b))
        return a.value, b.value

    def iterate(self, n_iterations, seed=None):
        """-> (found, bNoMore, vbInliers, nInliers, Tcw). `seed`: srand(seed) first (the draws of rand_draws(seed, ...))."""
        if seed is not None:
            lib().ref_srand(int(seed))
        no_more, n_inl = C.c_int32(), C.c_int32()
        inl = np.zeros(max(self.N, 1), np.uint8)
        T = np.zeros(16, np.float32)
        rc = lib().ref_pnp_iterate(self._h, int(n_iterations), C.byref(no_more), _ptr(inl), C.byref(n_inl), _ptr(T))
        return rc, bool(no_more.value), inl[:self.N].astype(bool), n_inl.value, T.reshape(4, 4)

    @property
    def iterations(self):
        return lib().ref_pnp_iterations(self._h)


def cv_gemm(A, B, alpha=1.0, Cm=None, beta=0.0, flags=0):
    A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
    Cm = None if Cm is None else np.ascontiguousarray(Cm, np.float32)
    m = A.shape[1] if flags & 1 else A.shape[0]
    n = B.shape[0] if flags & 2 else B.shape[1]
    D = np.zeros((m, n), np.float32)
    lib().ref_cv_gemm_f32(_ptr(A), A.shape[0], A.shape[1], _ptr(B), B.shape[0], B.shape[1], float(alpha), _ptr(Cm),
                          0 if Cm is None else Cm.shape[0], 0 if Cm is None else Cm.shape[1], float(beta), int(flags), _ptr(D))
    return D


def expr_minus_Rt_t(R, t):
    R = np.ascontiguousarray(R, np.float32); t = np.ascontiguousarray(t, np.float32); o = np.zeros(3, np.float32)
    lib().ref_expr_minus_Rt_t(_ptr(R), _ptr(t), _ptr(o))
    return o


def expr_Rx_plus_t(R, x, t):
    R = np.ascontiguousarray(R, np.float32); x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
    o = np.zeros(3, np.float32)
    lib().ref_expr_Rx_plus_t(_ptr(R), _ptr(x), _ptr(t), _ptr(o))
    return o
