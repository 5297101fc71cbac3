"""Host-side mirror of ORB_SLAM2::ORBextractor (reference: corbslam_client/include/ORBextractor.h:45-112) over the
C ABI of libcorb_b200.so. Same constructor arguments, same call semantics, same getters; `mvImagePyramid` is filled
on request (the reference exposes it as a public member that Frame::ComputeStereoMatches reads, Frame.cc:477).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, lib


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0):
        h = C.c_void_p()
        check(lib().corb_orb_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST),
                                    int(device), C.byref(h)))
        self._h = h
        self.nlevels = int(nlevels)
        n = self.nlevels
        self._scale = np.empty(n, np.float32)
        self._inv_scale = np.empty(n, np.float32)
        self._sigma2 = np.empty(n, np.float32)
        self._inv_sigma2 = np.empty(n, np.float32)
        self.mnFeaturesPerLevel = np.empty(n, np.int32)
        self.umax = np.empty(16, np.int32)
        p = lambda a, t: a.ctypes.data_as(t)
        check(lib().corb_orb_tables(h, p(self._scale, _lib.f32p), p(self._inv_scale, _lib.f32p), p(self._sigma2, _lib.f32p),
                                    p(self._inv_sigma2, _lib.f32p), p(self.mnFeaturesPerLevel, _lib.i32p),
                                    p(self.umax, _lib.i32p)))
        self.mvImagePyramid = []
        self._out_shape = None
        self._shape_v = None
        self.copy_outputs = True  # False: return views into reused buffers (valid until the next call)

    def close(self):
        if getattr(self, "_h", None):
            lib().corb_orb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals are already gone, the process frees the device memory
            pass

    @property
    def _shape(self):
        return self._shape_v

    @_shape.setter
    def _shape(self, v):
        if v != self._shape_v:  # a different image size rebuilds the plan: the page-locked result buffer moves
            self._hv_shape = None
        self._shape_v = v

    def _host_view(self, n, shape):
        """Slices of numpy views over the whole page-locked result buffer; the views are built once per plan (the buffer
        moves only when the image size changes), so a call costs two slices instead of a C call and two new arrays."""
        if n == 0:
            return np.empty(0, KP_DTYPE), None
        if getattr(self, "_hv_shape", None) != shape:
            kp, dp, cnt = C.c_void_p(), C.c_void_p(), C.c_int32()
            check(lib().corb_orb_host_results(self._h, C.byref(kp), C.byref(dp), C.byref(cnt)))
            cap = self.capacity(shape[1], shape[0])
            self._hv = (np.frombuffer((C.c_uint8 * (cap * KP_DTYPE.itemsize)).from_address(kp.value), KP_DTYPE),
                        np.frombuffer((C.c_uint8 * (cap * 32)).from_address(dp.value), np.uint8).reshape(cap, 32))
            self._hv_shape = shape
        return self._hv[0][:n], self._hv[1][:n]

    def host_results(self):
        """(keypoints, descriptors) of the last completed extraction as views of the handle's page-locked result buffer
        (corb_orb_host_results); valid until the next extraction on this handle."""
        kp, dp, n = C.c_void_p(), C.c_void_p(), C.c_int32()
        check(lib().corb_orb_host_results(self._h, C.byref(kp), C.byref(dp), C.byref(n)))
        if n.value == 0:
            return np.empty(0, KP_DTYPE), None
        k = np.frombuffer((C.c_uint8 * (n.value * KP_DTYPE.itemsize)).from_address(kp.value), KP_DTYPE)
        d = np.frombuffer((C.c_uint8 * (n.value * 32)).from_address(dp.value), np.uint8).reshape(n.value, 32)
        return k, d

    # ---- getters (ORBextractor.h:63-85)
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return lib().corb_orb_scale_factor(self._h)

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv_scale.copy()

    def GetScaleSigmaSquares(self):
        return self._sigma2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._inv_sigma2.copy()

    def level_size(self, level, w, h):
        lw, lh = C.c_int32(), C.c_int32()
        check(lib().corb_orb_level_size(self._h, level, w, h, C.byref(lw), C.byref(lh)))
        return lw.value, lh.value

    def capacity(self, w, h):
        return lib().corb_orb_capacity(self._h, w, h)

    # ---- operator() (ORBextractor.cc:1043-1105)
    def submit(self, image, want_pyramid=False):
        if image is None or image.size == 0:
            self._shape = None
            check(lib().corb_orb_extract_submit(self._h, None, 0, 0, 0, 0))
            return
        if image.dtype != np.uint8 or image.ndim != 2:
            raise TypeError("image must be a 2-D uint8 array (CV_8UC1)")  # the reference asserts (:1050)
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        self._keep = image
        self._shape = image.shape
        self._want_pyr = bool(want_pyramid)
        check(lib().corb_orb_extract_submit(self._h, image.ctypes.data, image.shape[1], image.shape[0], image.strides[0],
                                            int(self._want_pyr)))

    def wait(self):
        if self._shape is None:
            n = C.c_int32()
            check(lib().corb_orb_extract_wait(self._h, None, None, C.byref(n), None))
            self.mvImagePyramid = []
            return np.zeros(0, KP_DTYPE), None  # descriptors.release() (:1064-1065)
        h, w = self._shape
        if self._out_shape != (h, w):  # output buffers are reused between calls of the same size
            cap = self.capacity(w, h)
            self._out = (np.empty(cap, KP_DTYPE), np.empty((cap, 32), np.uint8))
            self._out_shape = (h, w)
        kps, desc = self._out
        n = C.c_int32()
        pyr_ptrs = None
        if self._want_pyr:
            pyr = [np.empty(self.level_size(l, w, h)[::-1], np.uint8) for l in range(self.nlevels)]
            pyr_ptrs = (C.c_void_p * self.nlevels)(*[a.ctypes.data for a in pyr])
        check(lib().corb_orb_extract_wait(self._h, kps.ctypes.data, desc.ctypes.data, C.byref(n), pyr_ptrs))
        if self._want_pyr:
            self.mvImagePyramid = pyr
        self._keep = None
        if n.value == 0:
            return kps[:0].copy(), None
        if self.copy_outputs:
            return kps[:n.value].copy(), desc[:n.value].copy()
        return kps[:n.value], desc[:n.value]  # views into buffers that the next call overwrites

    def __call__(self, image, mask=None, want_pyramid=False):
        """Returns (keypoints, descriptors): a structured array with cv::KeyPoint's fields and an (N,32) uint8 array
        (None when no keypoint was found, mirroring `_descriptors.release()`). `mask` is ignored like in the reference."""
        self.submit(image, want_pyramid)
        return self.wait()

    # ---- device-resident path (inputs already in HBM)
    def extract_device(self, d_ptr, w, h, stride):
        self._hv_shape = None  # (may rebuild the plan for another size)
        check(lib().corb_orb_extract_device(self._h, int(d_ptr), w, h, stride))

    def sync(self):
        check(lib().corb_orb_sync(self._h))

    def stream(self):
        return lib().corb_orb_stream(self._h)

    def uses_tma(self):
        return bool(lib().corb_orb_uses_tma(self._h))

    def set_host_transfer(self, mode):
        """corb_orb_set_host_transfer: 0 = the import kernel reads page-locked images in place (lowest latency of one blocking
        frame, default), 1 = copy-engine memcpy nodes (highest throughput with frames in flight). Set on both handles of a pair."""
        check(lib().corb_orb_set_host_transfer(self._h, int(mode)))
        return self

    def host_transfer(self):
        return int(lib().corb_orb_host_transfer(self._h))

    def launches_per_extract(self):
        return lib().corb_orb_launches_per_extract(self._h)

    def device_results(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().corb_orb_device_results(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def profile(self, reps=10):
        """[(kernel name, ms)] per launch of one extraction, CUDA-event timed (eager replay, not the graph)."""
        ms = np.zeros(64, np.float32)
        n = C.c_int32()
        check(lib().corb_orb_profile(self._h, reps, ms.ctypes.data_as(_lib.f32p), 64, C.byref(n)))
        return [(lib().corb_orb_kernel_name(self._h, i).decode(), float(ms[i])) for i in range(n.value)]

    # ---- stage taps for parity tests
    def tap_image(self, level, blurred=False):
        h, w = self._last_shape()
        lw, lh = self.level_size(level, w, h)
        out = np.empty((lh, lw), np.uint8)
        check(lib().corb_orb_tap(self._h, 1 if blurred else 0, level, out.ctypes.data, out.nbytes, None))
        return out

    def tap_candidates(self, level):
        n = C.c_int32()
        check(lib().corb_orb_tap(self._h, 2, level, None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 3), np.int32)
        check(lib().corb_orb_tap(self._h, 2, level, out.ctypes.data, out.nbytes, C.byref(n)))
        return out[:n.value]

    def tap_level_count(self, level):
        n = C.c_int32()
        check(lib().corb_orb_tap(self._h, 3, level, None, 0, C.byref(n)))
        return n.value

    def _last_shape(self):
        if getattr(self, "_shape", None) is None:
            raise RuntimeError("no extraction has run")
        return self._shape


def extract_stereo(ex_left, ex_right, left, right, want_pyramid=False):
    """Left/right extraction of one stereo frame from one thread (Frame::Frame's two ExtractORB threads, Frame.cc:78-81):
    -> ((kps_l, desc_l), (kps_r, desc_r))."""
    if left.shape != right.shape or left.dtype != np.uint8 or right.dtype != np.uint8 or left.ndim != 2:
        raise TypeError("left/right must be 2-D uint8 arrays of the same shape")
    if left.strides != right.strides or left.strides[1] != 1:
        left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
    h, w = left.shape
    outs = []
    for ex in (ex_left, ex_right):
        if ex._out_shape != (h, w):
            cap = ex.capacity(w, h)
            ex._out = (np.empty(cap, KP_DTYPE), np.empty((cap, 32), np.uint8))
            ex._out_shape = (h, w)
        ex._shape = (h, w)
        outs.append(ex._out)
    nl, nr = C.c_int32(), C.c_int32()
    pl = pr = None
    if not want_pyramid and not ex_left.copy_outputs and not ex_right.copy_outputs:
        # zero-copy: the results are read in place from the handles' page-locked result buffers
        check(lib().corb_orb_extract_pair(ex_left._h, ex_right._h, left.ctypes.data, right.ctypes.data, w, h, left.strides[0],
                                          None, None, C.byref(nl), None, None, C.byref(nr), None, None))
        return ex_left._host_view(nl.value, (h, w)), ex_right._host_view(nr.value, (h, w))
    if want_pyramid:
        pyrs = [[np.empty(ex.level_size(l, w, h)[::-1], np.uint8) for l in range(ex.nlevels)] for ex in (ex_left, ex_right)]
        pl = (C.c_void_p * ex_left.nlevels)(*[a.ctypes.data for a in pyrs[0]])
        pr = (C.c_void_p * ex_right.nlevels)(*[a.ctypes.data for a in pyrs[1]])
        ex_left.mvImagePyramid, ex_right.mvImagePyramid = pyrs
    check(lib().corb_orb_extract_pair(ex_left._h, ex_right._h, left.ctypes.data, right.ctypes.data, w, h, left.strides[0],
                                      outs[0][0].ctypes.data, outs[0][1].ctypes.data, C.byref(nl), outs[1][0].ctypes.data,
                                      outs[1][1].ctypes.data, C.byref(nr), pl, pr))
    res = []
    for (k, d), n, ex in zip(outs, (nl.value, nr.value), (ex_left, ex_right)):
        if n == 0:
            res.append((k[:0].copy(), None))
        elif ex.copy_outputs:
            res.append((k[:n].copy(), d[:n].copy()))
        else:
            res.append((k[:n], d[:n]))
    return tuple(res)


def extract_stereo_submit(ex_left, ex_right, left, right):
    """corb_orb_extract_pair_submit: enqueue one stereo pair (page-locked or pageable host images) and return. One client
    thread keeps two frames in flight by alternating between two handle pairs; extract_stereo_wait collects."""
    h, w = left.shape
    ex_left._shape = ex_right._shape = (h, w)
    ex_left._keep_pair = (left, right)
    check(lib().corb_orb_extract_pair_submit(ex_left._h, ex_right._h, left.ctypes.data, right.ctypes.data, w, h, left.strides[0], 0))


def extract_stereo_wait(ex_left, ex_right):
    """-> ((kps_l, desc_l), (kps_r, desc_r)) as views of the handles' page-locked result buffers (valid until the next
    extraction on these handles)."""
    nl, nr = C.c_int32(), C.c_int32()
    check(lib().corb_orb_extract_pair_wait(ex_left._h, ex_right._h, None, None, C.byref(nl), None, None, C.byref(nr), None, None))
    ex_left._keep_pair = None
    return ex_left._host_view(nl.value, ex_left._shape), ex_right._host_view(nr.value, ex_right._shape)


def frame_stereo_submit(ex_left, ex_right, left, right, mbf, mb):
    """corb_frame_stereo_submit (ExtractORB x2 + ComputeStereoMatches enqueued as one graph launch)."""
    h, w = left.shape
    ex_left._shape = ex_right._shape = (h, w)
    ex_left._keep_pair = (left, right)
    check(lib().corb_frame_stereo_submit(ex_left._h, ex_right._h, left.ctypes.data, right.ctypes.data, w, h, left.strides[0],
                                         float(mbf), float(mb)))


def frame_stereo_wait(ex_left, ex_right):
    """-> ((kps_l, desc_l), (kps_r, desc_r), mvuRight, mvDepth); keypoints / descriptors are views of the result buffers."""
    cap = ex_left.capacity(ex_left._shape[1], ex_left._shape[0])
    if getattr(ex_left, "_stereo_out", None) is None or len(ex_left._stereo_out[0]) != cap:
        ex_left._stereo_out = (np.empty(cap, np.float32), np.empty(cap, np.float32))
    ur, dp = ex_left._stereo_out
    nl, nr = C.c_int32(), C.c_int32()
    check(lib().corb_frame_stereo_wait(ex_left._h, ex_right._h, None, None, C.byref(nl), None, None, C.byref(nr), ur.ctypes.data,
                                       dp.ctypes.data))
    ex_left._keep_pair = None
    return (ex_left._host_view(nl.value, ex_left._shape), ex_right._host_view(nr.value, ex_right._shape), ur[:nl.value],
            dp[:nl.value])


def extract_stereo_device(ex_left, ex_right, d_left, d_right, w, h, stride):
    if (h, w) != ex_left._shape_v or (h, w) != ex_right._shape_v:
        ex_left._hv_shape = ex_right._hv_shape = None  # another size rebuilds the plans
    check(lib().corb_orb_extract_pair_device(ex_left._h, ex_right._h, int(d_left), int(d_right), w, h, stride))


def compute_stereo_matches(ex_left, ex_right, n_left, mbf, mb):
    """Frame::ComputeStereoMatches (Frame.cc:470-644) on the last extraction of the two handles -> (mvuRight, mvDepth)."""
    ur = np.empty(max(n_left, 1), np.float32)
    dp = np.empty(max(n_left, 1), np.float32)
    check(lib().corb_stereo_match(ex_left._h, ex_right._h, float(mbf), float(mb), int(n_left), ur.ctypes.data, dp.ctypes.data))
    return ur[:n_left], dp[:n_left]


def frame_stereo(ex_left, ex_right, left, right, mbf, mb):
    """ExtractORB left + right and ComputeStereoMatches in one call (the stereo Frame constructor, Frame.cc:61-117):
    -> ((kps_l, desc_l), (kps_r, desc_r), mvuRight, mvDepth)."""
    if left.shape != right.shape or left.dtype != np.uint8 or right.dtype != np.uint8 or left.ndim != 2:
        raise TypeError("left/right must be 2-D uint8 arrays of the same shape")
    if left.strides != right.strides or left.strides[1] != 1:
        left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
    h, w = left.shape
    outs = []
    for ex in (ex_left, ex_right):
        if ex._out_shape != (h, w):
            cap = ex.capacity(w, h)
            ex._out = (np.empty(cap, KP_DTYPE), np.empty((cap, 32), np.uint8))
            ex._out_shape = (h, w)
        ex._shape = (h, w)
        outs.append(ex._out)
    if getattr(ex_left, "_stereo_out", None) is None or len(ex_left._stereo_out[0]) != len(outs[0][0]):
        ex_left._stereo_out = (np.empty(len(outs[0][0]), np.float32), np.empty(len(outs[0][0]), np.float32))
    ur, dp = ex_left._stereo_out
    nl, nr = C.c_int32(), C.c_int32()
    if not ex_left.copy_outputs and not ex_right.copy_outputs:  # zero-copy keypoints / descriptors (corb_orb_host_results)
        check(lib().corb_frame_stereo(ex_left._h, ex_right._h, left.ctypes.data, right.ctypes.data, w, h, left.strides[0], float(mbf),
                                      float(mb), None, None, C.byref(nl), None, None, C.byref(nr), ur.ctypes.data, dp.ctypes.data))
        return ex_left._host_view(nl.value, (h, w)), ex_right._host_view(nr.value, (h, w)), ur[:nl.value], dp[:nl.value]
    check(lib().corb_frame_stereo(ex_left._h, ex_right._h, left.ctypes.data, right.ctypes.data, w, h, left.strides[0], float(mbf),
                                  float(mb), outs[0][0].ctypes.data, outs[0][1].ctypes.data, C.byref(nl), outs[1][0].ctypes.data,
                                  outs[1][1].ctypes.data, C.byref(nr), ur.ctypes.data, dp.ctypes.data))
    cp = (lambda a: a.copy()) if ex_left.copy_outputs else (lambda a: a)
    return ((cp(outs[0][0][:nl.value]), cp(outs[0][1][:nl.value])), (cp(outs[1][0][:nr.value]), cp(outs[1][1][:nr.value])),
            cp(ur[:nl.value]), cp(dp[:nl.value]))
