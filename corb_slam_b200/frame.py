"""Flattened view of the reference's Frame for the projection matchers (reference: corbslam_client/include/Frame.h:38-39,
60-200; src/Frame.cc:229-245 AssignFeaturesToGrid, :386-397 PosInGrid). The C++ shim fills `corb_frame_view` from a Frame
object; this module builds the same struct from numpy arrays for the Python host layer and the tests."""
import ctypes as C

import numpy as np

FRAME_GRID_ROWS, FRAME_GRID_COLS = 48, 64  # Frame.h:38-39

_f32p, _i32p, _u8p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)


class FrameViewStruct(C.Structure):
    """corb_frame_view (include/corb_b200.h)."""
    _fields_ = [("n", C.c_int32), ("x", _f32p), ("y", _f32p), ("octave", _i32p), ("angle", _f32p), ("desc", _u8p),
                ("u_right", _f32p), ("taken", _u8p), ("grid_off", _i32p), ("grid_idx", _i32p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float), ("scale_factors", _f32p), ("n_levels", C.c_int32),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float),
                ("mb", C.c_float), ("Tcw", C.c_float * 12)]


def round_half_away(v):
    """C round() on float32 values."""
    v = np.asarray(v, np.float64)
    return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int64)


class FrameView:
    """What SearchByProjection reads from a Frame: undistorted keypoints, descriptors, mvuRight, the feature grid,
    image bounds, scale factors, intrinsics and the pose."""

    def __init__(self, x, y, octave, angle, desc, u_right, scale_factors, bounds, K, mbf, Tcw, taken=None):
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        self.x, self.y, self.angle, self.u_right = f32(x), f32(y), f32(angle), f32(u_right)
        self.octave = np.ascontiguousarray(octave, np.int32)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.n = len(self.x)
        self.scale_factors = f32(scale_factors)
        self.min_x, self.min_y, self.max_x, self.max_y = [np.float32(b) for b in bounds]
        self.fx, self.fy, self.cx, self.cy = [np.float32(k) for k in K]
        self.mbf = np.float32(mbf)
        self.mb = np.float32(self.mbf / self.fx)  # Frame.cc:225
        self.Tcw = f32(Tcw).reshape(-1)[:12].copy()
        self.taken = None if taken is None else np.ascontiguousarray(taken, np.uint8)
        # static members mfGridElementWidthInv / HeightInv (Frame.cc:101-102)
        self.grid_w_inv = np.float32(np.float32(FRAME_GRID_COLS) / np.float32(self.max_x - self.min_x))
        self.grid_h_inv = np.float32(np.float32(FRAME_GRID_ROWS) / np.float32(self.max_y - self.min_y))
        self.grid_off, self.grid_idx = self._assign_features_to_grid()

    def _assign_features_to_grid(self):
        """Frame::AssignFeaturesToGrid + PosInGrid: cell = round((pt - min) * inv), features outside the grid are
        dropped, push_back order = ascending feature index. CSR with cell = ix * 48 + iy."""
        px = round_half_away((self.x - self.min_x) * self.grid_w_inv)
        py = round_half_away((self.y - self.min_y) * self.grid_h_inv)
        ok = (px >= 0) & (px < FRAME_GRID_COLS) & (py >= 0) & (py < FRAME_GRID_ROWS)
        cell = (px * FRAME_GRID_ROWS + py)[ok]
        idx = np.nonzero(ok)[0]
        order = np.argsort(cell, kind="stable")
        off = np.zeros(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, np.int32)
        np.add.at(off, cell + 1, 1)
        return np.cumsum(off).astype(np.int32), np.ascontiguousarray(idx[order], np.int32)

    def c_struct(self):
        s = FrameViewStruct()
        p = lambda a, t: a.ctypes.data_as(t)
        s.n = self.n
        s.x, s.y, s.angle, s.u_right = p(self.x, _f32p), p(self.y, _f32p), p(self.angle, _f32p), p(self.u_right, _f32p)
        s.octave, s.desc = p(self.octave, _i32p), p(self.desc, _u8p)
        s.taken = p(self.taken, _u8p) if self.taken is not None else None
        s.grid_off, s.grid_idx = p(self.grid_off, _i32p), p(self.grid_idx, _i32p)
        s.min_x, s.min_y, s.max_x, s.max_y = self.min_x, self.min_y, self.max_x, self.max_y
        s.grid_w_inv, s.grid_h_inv = self.grid_w_inv, self.grid_h_inv
        s.scale_factors, s.n_levels = p(self.scale_factors, _f32p), len(self.scale_factors)
        s.fx, s.fy, s.cx, s.cy, s.mbf, s.mb = self.fx, self.fy, self.cx, self.cy, self.mbf, self.mb
        for i in range(12):
            s.Tcw[i] = float(self.Tcw[i])
        return s
