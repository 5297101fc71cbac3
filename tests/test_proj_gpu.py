"""GPU parity of ORBmatcher::SearchByProjection (both variants, ORBmatcher.cc:44-131, 1470-1614) against the CPU oracle:
bit-exact match arrays and counts on seeded synthetic scenes, including the order-dependent greedy semantics
(features taken by earlier queries), temporal (non-blocking) MapPoints, the rotation histogram and forward /
backward level windows."""
import numpy as np
import pytest

import oracle
from oracle import _proj_bind as PB
from corb_slam_b200 import FrameView, ORBmatcher
from corb_slam_b200.synth import projection_scene

pytestmark = pytest.mark.gpu


def _frame(s, Tcw=None, taken=True):
    c = s["cur"]
    return FrameView(c["x"], c["y"], c["octave"], c["angle"], c["desc"], c["u_right"], s["scales"], s["bounds"], s["K"], s["mbf"],
                     s["Tcw"] if Tcw is None else Tcw, taken=s["taken"] if taken else None)


@pytest.mark.parametrize("seed,th,mono,check_ori,blocks", [(1, 15.0, False, True, True), (2, 7.0, False, True, False),
                                                          (3, 15.0, True, False, True), (4, 30.0, False, True, True)])
def test_last_frame_variant_bit_exact(seed, th, mono, check_ori, blocks):
    oracle.lib()
    s = projection_scene(seed)
    fv = _frame(s)
    lb = s["last_blocks"] if blocks else None
    om, on = PB.search_by_projection_last(fv.c_struct(), fv.n, s["last_valid"], lb, s["Xw"], s["mp_desc"], s["last"]["octave"],
                                          s["last"]["angle"], s["Tlw"], th, mono, check_ori)
    m = ORBmatcher(0.9, check_ori)
    gm, gn = m.SearchByProjectionLastFrame(fv, s["last_valid"], s["Xw"], s["mp_desc"], s["last"]["octave"], s["last"]["angle"], s["Tlw"],
                                           th, bMono=mono, last_blocks=lb)
    assert gn == on and on > 300
    np.testing.assert_array_equal(gm, om)
    m.close()


def test_last_frame_forward_backward_and_no_motion():
    """tlc.z > mb selects the forward window (levels >= last octave), -tlc.z > mb the backward one (:1490-1491, 1528-1533)."""
    oracle.lib()
    m = ORBmatcher(0.9, True)
    for motion in (-2.0, 0.0, 2.0):
        s = projection_scene(6, motion=abs(motion) if motion else 0.05)
        Tcw = s["Tcw"].copy()
        if motion < 0:
            Tcw[2, 3] = -Tcw[2, 3]
        fv = _frame(s, Tcw=Tcw)
        om, on = PB.search_by_projection_last(fv.c_struct(), fv.n, s["last_valid"], s["last_blocks"], s["Xw"], s["mp_desc"],
                                              s["last"]["octave"], s["last"]["angle"], s["Tlw"], 15.0, False, True)
        gm, gn = m.SearchByProjectionLastFrame(fv, s["last_valid"], s["Xw"], s["mp_desc"], s["last"]["octave"], s["last"]["angle"],
                                               s["Tlw"], 15.0, last_blocks=s["last_blocks"])
        assert gn == on
        np.testing.assert_array_equal(gm, om)
    m.close()


@pytest.mark.parametrize("seed,th,nnratio", [(1, 1.0, 0.8), (2, 3.0, 0.8), (5, 5.0, 0.6)])
def test_map_point_variant_bit_exact(seed, th, nnratio):
    oracle.lib()
    s = projection_scene(seed)
    fv = _frame(s)
    om, on = PB.search_by_projection_map(fv.c_struct(), fv.n, s["in_view"], None, s["proj"], s["level"], s["view_cos"], s["mp_desc"], th,
                                         nnratio)
    m = ORBmatcher(nnratio, True)
    gm, gn = m.SearchByProjectionMapPoints(fv, s["in_view"], s["proj"], s["level"], s["view_cos"], s["mp_desc"], th=th)
    assert gn == on and on > 500
    np.testing.assert_array_equal(gm, om)
    m.close()


def test_dense_windows_and_degenerate_inputs():
    """A crowded image region overflows the first candidate buffer (128 per query) and takes the retry; empty inputs."""
    oracle.lib()
    s = projection_scene(9, n_points=6000, clutter=3000, w=400, h=300)
    fv = _frame(s)
    om, on = PB.search_by_projection_last(fv.c_struct(), fv.n, s["last_valid"], s["last_blocks"], s["Xw"], s["mp_desc"],
                                          s["last"]["octave"], s["last"]["angle"], s["Tlw"], 40.0, False, True)
    m = ORBmatcher(0.9, True)
    gm, gn = m.SearchByProjectionLastFrame(fv, s["last_valid"], s["Xw"], s["mp_desc"], s["last"]["octave"], s["last"]["angle"], s["Tlw"],
                                           40.0, last_blocks=s["last_blocks"])
    assert gn == on
    np.testing.assert_array_equal(gm, om)
    none = np.zeros(len(s["last_valid"]), np.uint8)
    gm, gn = m.SearchByProjectionLastFrame(fv, none, s["Xw"], s["mp_desc"], s["last"]["octave"], s["last"]["angle"], s["Tlw"], 15.0)
    assert gn == 0 and np.all(gm == -1)
    m.close()
