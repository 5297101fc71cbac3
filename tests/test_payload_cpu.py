"""Binary KeyFrame payload (corb_kf_payload_*): bit-exact round trip of the extractor-produced
KeyFrame members, rejection of foreign / damaged input, and its size beside the decimal text the reference's boost text
archive writes for the same members (KeyFrame.h:61-87, SerializeObject.h:34-61). Host code: runs without a GPU."""
import numpy as np
import pytest

import oracle
from oracle import _match_bind as M
from corb_slam_b200 import keyframe_payload as KP
from corb_slam_b200._lib import CorbError
from corb_slam_b200.synth import stereo_frame


def _frame(seed=1234, w=640, h=240):
    L, R = stereo_frame(seed, w=w, h=h)
    exl, exr = oracle.OrbExtractor(800, 1.2, 6, 20, 7), oracle.OrbExtractor(800, 1.2, 6, 20, 7)
    kl, dl = exl(L)
    kr, dr = exr(R)
    ur, dp, _ = oracle.stereo_matches(exl, exr, kl, dl, kr, dr, 386.1448, np.float32(386.1448) / np.float32(718.856))
    voc = M.Vocabulary.from_arrays(*M.random_vocabulary(10, 4, 3))
    b = voc.transform(dl, 2)
    return kl, ur, dp, dl, (b[0], b[1]), (b[2], b[3], b[4])


def _text_archive_chars(keys, keys_un, ur, dp, desc, bow, fv):
    """Characters a text archive needs for the same members: every number in decimal (floats with 9 significant digits, the
    precision boost uses for float; doubles with 17), separated by one space - a lower bound (no class / version tokens)."""
    f = lambda a: sum(len("%.9g" % v) + 1 for v in np.asarray(a, np.float64).ravel())
    i = lambda a: sum(len(str(int(v))) + 1 for v in np.asarray(a).ravel())
    n = 0
    for k in (keys, keys_un):
        n += f(k["x"]) + f(k["y"]) + f(k["angle"]) + f(k["size"]) + f(k["response"]) + i(k["octave"]) + i(k["class_id"])
    n += f(ur) + f(dp) + i(desc)
    n += i(bow[0]) + sum(len("%.17g" % v) + 1 for v in bow[1]) + i(fv[0]) + i(fv[2]) + 2 * len(fv[0])
    return n


def test_round_trip_is_bit_exact_and_three_times_smaller_than_text():
    oracle.lib()
    keys, ur, dp, desc, bow, fv = _frame()
    blob = KP.encode(keys, None, ur, dp, desc, bow, fv)
    out = KP.decode(blob)
    assert out["same_un"] and out["keys"].tobytes() == keys.tobytes() and out["keys_un"].tobytes() == keys.tobytes()
    assert out["u_right"].tobytes() == ur.tobytes() and out["depth"].tobytes() == dp.tobytes() and np.array_equal(out["desc"], desc)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(out["bow"] + out["fv"], bow + fv))
    per_kp = len(blob) / len(keys)
    text = _text_archive_chars(keys, keys, ur, dp, desc, bow, fv)
    print("payload %d B (%.1f B / keypoint), text archive >= %d chars (%.1f / keypoint)" % (len(blob), per_kp, text, text / len(keys)))
    assert per_kp < 80 and text > 3.0 * len(blob)
    # undistorted keypoints that differ (RGB-D with lens distortion) and class ids are kept
    un = keys.copy()
    un["x"] += np.float32(0.25)
    un["class_id"][3] = 7
    out = KP.decode(KP.encode(keys, un, ur, dp, desc, bow, fv))
    assert not out["same_un"] and out["keys_un"].tobytes() == un.tobytes() and out["keys"].tobytes() == keys.tobytes()


def test_empty_and_foreign_and_damaged_payloads():
    e = np.zeros(0, KP.KP_DTYPE)
    z32, zu = np.zeros(0, np.float32), np.zeros(0, np.uint32)
    blob = KP.encode(e, None, z32, z32, np.zeros((0, 32), np.uint8), (zu, np.zeros(0)), (zu, np.zeros(1, np.int32), zu))
    out = KP.decode(blob)
    assert len(out["keys"]) == 0 and len(out["bow"][0]) == 0 and len(blob) == 40
    oracle.lib()
    keys, ur, dp, desc, bow, fv = _frame(7, 320, 240)
    blob = bytearray(KP.encode(keys, None, ur, dp, desc, bow, fv))
    with pytest.raises(CorbError):
        KP.decode(b"22 serialization::archive 10 0 0 0 0 " + bytes(blob))  # what a boost text archive starts with
    with pytest.raises(CorbError):
        KP.decode(bytes(blob[:len(blob) // 2]))                               # truncated
    blob[100] ^= 0x40
    with pytest.raises(CorbError):
        KP.decode(bytes(blob))                                                # checksum
