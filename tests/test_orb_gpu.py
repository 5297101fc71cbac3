"""GPU parity of the ORB front end against the CPU oracle, stage by stage and end to end (bit-exact)."""
import numpy as np
import pytest

import oracle
from corb_slam_b200 import ORBextractor
from corb_slam_b200.synth import stereo_frame

pytestmark = pytest.mark.gpu

PARAMS = (2000, 1.2, 8, 20, 7)  # KITTI00-02.yaml:38-51


def _compare(img, params=PARAMS, stages=True):
    ora = oracle.OrbExtractor(*params)
    gpu = ORBextractor(*params)
    okp, odesc = ora(img)
    gkp, gdesc = gpu(img, want_pyramid=True)
    h, w = img.shape
    if stages:
        for l in range(params[2]):
            np.testing.assert_array_equal(gpu.tap_image(l), ora.pyramid(l), err_msg="pyramid level %d" % l)
            np.testing.assert_array_equal(gpu.mvImagePyramid[l], ora.pyramid(l), err_msg="host pyramid level %d" % l)
            np.testing.assert_array_equal(gpu.tap_candidates(l), ora.candidates(l), err_msg="candidates level %d" % l)
            assert gpu.tap_level_count(l) == ora.level_count(l), "kept keypoints level %d" % l
            ob = ora.blurred(l)
            if ob is not None:
                np.testing.assert_array_equal(gpu.tap_image(l, blurred=True), ob, err_msg="blurred level %d" % l)
    assert len(gkp) == len(okp)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        np.testing.assert_array_equal(gkp[f], okp[f], err_msg=f)
    np.testing.assert_array_equal(gkp["angle"].view(np.uint32), okp["angle"].view(np.uint32), err_msg="angle bits")
    if len(okp):
        np.testing.assert_array_equal(gdesc, odesc)
    else:
        assert gdesc is None
    gpu.close()
    return len(gkp)


@pytest.mark.parametrize("seed", [1234, 1235, 1236])
def test_synthetic_frames_bit_exact(seed):
    left, right = stereo_frame(seed)
    assert _compare(left) >= 2000
    assert _compare(right, stages=False) >= 2000


@pytest.mark.parametrize("size", [(1241, 376), (1226, 370), (640, 480), (752, 480)])
def test_other_sizes(size):
    left, _ = stereo_frame(7, w=size[0], h=size[1])
    _compare(left)


def test_noise_image_many_candidates():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (375, 1242), dtype=np.uint8)
    _compare(img)


def test_flat_image_no_keypoints():
    img = np.full((375, 1242), 77, np.uint8)
    assert _compare(img) == 0


def test_low_texture_threshold_fallback():
    left, _ = stereo_frame(11)
    img = (left.astype(np.int32) // 6 + 100).astype(np.uint8)  # contrast low enough that most cells need minThFAST
    _compare(img)


def test_few_features_and_levels():
    left, _ = stereo_frame(5, w=640, h=480)
    _compare(left, params=(300, 1.2, 4, 20, 7))
    _compare(left, params=(1000, 1.5, 3, 12, 5))


def test_empty_image_returns_nothing():
    gpu = ORBextractor(*PARAMS)
    kps, desc = gpu(np.zeros((0, 0), np.uint8))
    assert len(kps) == 0 and desc is None
    gpu.close()


def test_strided_input_and_two_handles_concurrently():
    import threading
    left, right = stereo_frame(1240)
    big = np.zeros((375, 1300), np.uint8)
    big[:, :1242] = left
    view = big[:, :1242]
    ora = oracle.OrbExtractor(*PARAMS)
    okl = ora(left)
    okr = ora(right)
    exl, exr = ORBextractor(*PARAMS), ORBextractor(*PARAMS)
    res = {}
    def run(name, ex, im):
        for _ in range(5):
            res[name] = ex(im)
    t1 = threading.Thread(target=run, args=("l", exl, view))
    t2 = threading.Thread(target=run, args=("r", exr, right))
    t1.start(); t2.start(); t1.join(); t2.join()
    for got, exp in ((res["l"], okl), (res["r"], okr)):
        assert got[0].tobytes() == exp[0].tobytes()
        np.testing.assert_array_equal(got[1], exp[1])
    exl.close(); exr.close()


def test_stereo_pair_call_and_device_resident_path():
    import torch
    from corb_slam_b200 import extract_stereo, extract_stereo_device
    left, right = stereo_frame(1250)
    ora = oracle.OrbExtractor(*PARAMS)
    okl, okr = ora(left), ora(right)
    exl, exr = ORBextractor(*PARAMS), ORBextractor(*PARAMS)
    # pageable numpy input and page-locked input (read in place over PCIe by the import kernel) must agree
    pl, pr = torch.from_numpy(left).pin_memory(), torch.from_numpy(right).pin_memory()
    for a, b in ((left, right), (pl.numpy(), pr.numpy())):
        (kl, dl), (kr, dr) = extract_stereo(exl, exr, a, b, want_pyramid=True)
        assert kl.tobytes() == okl[0].tobytes() and kr.tobytes() == okr[0].tobytes()
        np.testing.assert_array_equal(dl, okl[1]); np.testing.assert_array_equal(dr, okr[1])
        np.testing.assert_array_equal(exr.mvImagePyramid[3], ora.pyramid(3))
    # device-resident: pitched device image in, results stay in HBM
    dl_, dr_ = torch.zeros((375, 1280), dtype=torch.uint8, device="cuda"), torch.zeros((375, 1280), dtype=torch.uint8, device="cuda")
    dl_[:, :1242] = torch.from_numpy(left).cuda(); dr_[:, :1242] = torch.from_numpy(right).cuda()
    extract_stereo_device(exl, exr, dl_.data_ptr(), dr_.data_ptr(), 1242, 375, 1280)
    exl.sync(); exr.sync()
    kp_ptr, desc_ptr, cnt_ptr = exr.device_results()
    assert exr.tap_level_count(0) == ora.level_count(0)
    np.testing.assert_array_equal(exr.tap_image(2), ora.pyramid(2))
    exl.close(); exr.close()


def test_tma_and_fallback_staging_agree(monkeypatch):
    left, _ = stereo_frame(1260)
    ex = ORBextractor(*PARAMS)
    k1, d1 = ex(left)
    assert ex.uses_tma(), "cuTensorMapEncodeTiled must be available on the B200 box"
    ex.close()
    monkeypatch.setenv("CORB_NO_TMA", "1")
    ex2 = ORBextractor(*PARAMS)
    k2, d2 = ex2(left)
    assert not ex2.uses_tma()
    assert k1.tobytes() == k2.tobytes()
    np.testing.assert_array_equal(d1, d2)
    ex2.close()


KITTI_BF, KITTI_FX = 386.1448, 718.856


@pytest.mark.parametrize("seed,disparity", [(1234, 8), (1270, 23), (1271, 1), (1272, 0)])
def test_compute_stereo_matches_bit_exact(seed, disparity):
    from corb_slam_b200 import compute_stereo_matches, extract_stereo, frame_stereo
    left, right = stereo_frame(seed, disparity=disparity)
    mbf, mb = KITTI_BF, KITTI_BF / KITTI_FX
    ol, orr = oracle.OrbExtractor(*PARAMS), oracle.OrbExtractor(*PARAMS)
    (okl, odl), (okr, odr) = ol(left), orr(right)
    our, odp, kept = oracle.stereo_matches(ol, orr, okl, odl, okr, odr, mbf, mb)
    exl, exr = ORBextractor(*PARAMS), ORBextractor(*PARAMS)
    (kl, dl), (kr, dr) = extract_stereo(exl, exr, left, right)
    ur, dp = compute_stereo_matches(exl, exr, len(kl), mbf, mb)
    np.testing.assert_array_equal(ur.view(np.uint32), our.view(np.uint32))
    np.testing.assert_array_equal(dp.view(np.uint32), odp.view(np.uint32))
    assert int((ur >= 0).sum()) == kept and (kept > 800 or disparity == 0)
    # the one-call form (stereo Frame constructor) gives the same
    (kl2, dl2), (kr2, dr2), ur2, dp2 = frame_stereo(exl, exr, left, right, mbf, mb)
    assert kl2.tobytes() == okl.tobytes() and kr2.tobytes() == okr.tobytes()
    np.testing.assert_array_equal(ur2.view(np.uint32), our.view(np.uint32))
    np.testing.assert_array_equal(dp2.view(np.uint32), odp.view(np.uint32))
    exl.close(); exr.close()


def test_stereo_matches_unrelated_images():
    from corb_slam_b200 import frame_stereo
    left, _ = stereo_frame(1280)
    _, right = stereo_frame(1281)
    mbf, mb = KITTI_BF, KITTI_BF / KITTI_FX
    ol, orr = oracle.OrbExtractor(*PARAMS), oracle.OrbExtractor(*PARAMS)
    (okl, odl), (okr, odr) = ol(left), orr(right)
    our, odp, kept = oracle.stereo_matches(ol, orr, okl, odl, okr, odr, mbf, mb)
    exl, exr = ORBextractor(*PARAMS), ORBextractor(*PARAMS)
    _, _, ur, dp = frame_stereo(exl, exr, left, right, mbf, mb)
    np.testing.assert_array_equal(ur.view(np.uint32), our.view(np.uint32))
    np.testing.assert_array_equal(dp.view(np.uint32), odp.view(np.uint32))
    exl.close(); exr.close()


def test_quadtree_closed_form_equals_pass_by_pass(monkeypatch):
    """k_octtree's closed-form path (histogram of depth-5 cells) and its generic pass-by-pass code give the same
    keypoints in the same order; both equal the oracle (DistributeOctTree, ORBextractor.cc:539-763)."""
    for seed, size, params in [(21, (1242, 375), PARAMS), (22, (640, 480), (1000, 1.2, 8, 20, 7)),
                               (23, (1242, 375), (5000, 1.2, 8, 20, 7)), (24, (400, 300), (150, 1.3, 5, 20, 7))]:
        img, _ = stereo_frame(seed, w=size[0], h=size[1])
        monkeypatch.delenv("CORB_OCT_GENERIC", raising=False)
        fast = ORBextractor(*params)
        kf, df = fast(img)
        monkeypatch.setenv("CORB_OCT_GENERIC", "1")
        gen = ORBextractor(*params)
        kg, dg = gen(img)
        monkeypatch.delenv("CORB_OCT_GENERIC", raising=False)
        okp, odesc = oracle.OrbExtractor(*params)(img)
        assert kf.tobytes() == kg.tobytes() == okp.tobytes()
        np.testing.assert_array_equal(df, dg)
        np.testing.assert_array_equal(df, odesc)
        fast.close(); gen.close()


def test_sparse_and_clustered_keys_fall_back_to_generic_quadtree():
    """Few corners (fewer than the quota) and tight clusters need quadtree depths beyond the tabulated ones."""
    img = np.full((375, 1242), 60, np.uint8)
    rng = np.random.default_rng(5)
    for _ in range(150):  # isolated bright squares -> a few hundred corners, far fewer than 2000
        x, y = int(rng.integers(30, 1200)), int(rng.integers(30, 340))
        img[y:y + 5, x:x + 5] = 200
    _compare(img)
    img2 = np.full((375, 1242), 60, np.uint8)
    for _ in range(400):  # everything inside one 120 x 90 window
        x, y = int(rng.integers(600, 720)), int(rng.integers(150, 240))
        img2[y:y + 3, x:x + 3] = int(rng.integers(120, 255))
    _compare(img2)


def test_zero_copy_host_results_equal_copied_results():
    """corb_orb_host_results: NULL output pointers + views of the page-locked result buffer give the same keypoints and
    descriptors as the copying form (and as the oracle)."""
    from corb_slam_b200 import extract_stereo
    left, right = stereo_frame(31)
    a, b = ORBextractor(*PARAMS), ORBextractor(*PARAMS)
    (kl, dl), (kr, dr) = extract_stereo(a, b, left, right)
    a.copy_outputs = b.copy_outputs = False
    (vl, wl), (vr, wr) = extract_stereo(a, b, left, right)
    assert vl.tobytes() == kl.tobytes() and vr.tobytes() == kr.tobytes()
    np.testing.assert_array_equal(wl, dl)
    np.testing.assert_array_equal(wr, dr)
    okp, odesc = oracle.OrbExtractor(*PARAMS)(right)
    assert vr.tobytes() == okp.tobytes()
    np.testing.assert_array_equal(wr, odesc)
    a.close(); b.close()


def test_zero_copy_views_survive_a_size_change():
    """The cached views of the page-locked result buffer are rebuilt when the image size (and with it the plan) changes."""
    from corb_slam_b200 import extract_stereo
    a, b = ORBextractor(*PARAMS), ORBextractor(*PARAMS)
    a.copy_outputs = b.copy_outputs = False
    for seed, size in [(41, (1242, 375)), (42, (640, 480)), (43, (1242, 375))]:
        left, right = stereo_frame(seed, w=size[0], h=size[1])
        (vl, wl), (vr, wr) = extract_stereo(a, b, left, right)
        okp, odesc = oracle.OrbExtractor(*PARAMS)(left)
        assert vl.tobytes() == okp.tobytes()
        np.testing.assert_array_equal(wl, odesc)
    a.close(); b.close()


def test_repeated_and_concurrent_extraction_is_deterministic():
    """Size-independent properties at the bench size: the same frame gives the same bytes on every call (the FAST
    candidate lists are filled in arrival order, the quadtree must not depend on it), on both handles of a pair, and
    from several client threads sharing the GPU (per-handle streams, counters and graphs)."""
    import threading
    from corb_slam_b200 import extract_stereo
    left, right = stereo_frame(51)
    ref = ORBextractor(*PARAMS)
    k0, d0 = ref(left)
    k1, d1 = ref(right)
    for _ in range(10):
        k, d = ref(left)
        assert k.tobytes() == k0.tobytes() and d.tobytes() == d0.tobytes()
    errors = []

    def client(idx):
        try:
            a, b = ORBextractor(*PARAMS), ORBextractor(*PARAMS)
            for i in range(15):
                (kl, dl), (kr, dr) = extract_stereo(a, b, left, right) if (i + idx) % 2 else extract_stereo(b, a, left, right)
                assert kl.tobytes() == k0.tobytes() and dl.tobytes() == d0.tobytes()
                assert kr.tobytes() == k1.tobytes() and dr.tobytes() == d1.tobytes()
            a.close(); b.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ths = [threading.Thread(target=client, args=(i,)) for i in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errors, errors
    ref.close()


def test_two_stereo_frames_in_flight_from_one_thread():
    """corb_orb_extract_pair_submit / _wait and corb_frame_stereo_submit / _wait on two handle pairs: frame i + 1 is in flight
    while frame i is collected; results equal the blocking calls (and therefore the oracle) bit for bit."""
    from corb_slam_b200 import (extract_stereo, extract_stereo_submit, extract_stereo_wait, frame_stereo, frame_stereo_submit,
                                frame_stereo_wait)
    from corb_slam_b200 import _lib
    frames = [stereo_frame(1300 + i) for i in range(5)]
    pairs = [(ORBextractor(*PARAMS), ORBextractor(*PARAMS)) for _ in range(2)]
    ref_pair = (ORBextractor(*PARAMS), ORBextractor(*PARAMS))
    for a, b in pairs:
        a.copy_outputs = b.copy_outputs = False
    mbf, mb = 386.1448, np.float32(386.1448) / np.float32(718.856)
    want = [tuple((k.copy(), d.copy()) for k, d in extract_stereo(*ref_pair, l, r)) for l, r in frames]
    extract_stereo_submit(*pairs[0], *frames[0])
    for i in range(len(frames)):
        if i + 1 < len(frames):
            extract_stereo_submit(*pairs[(i + 1) % 2], *frames[i + 1])
        (kl, dl), (kr, dr) = extract_stereo_wait(*pairs[i % 2])
        assert kl.tobytes() == want[i][0][0].tobytes() and np.array_equal(dl, want[i][0][1])
        assert kr.tobytes() == want[i][1][0].tobytes() and np.array_equal(dr, want[i][1][1])
    wantf = []
    for l, r in frames:
        (kl, dl), (kr, dr), ur, dp = frame_stereo(*ref_pair, l, r, mbf, mb)
        wantf.append((kl.copy(), ur.copy(), dp.copy()))
    frame_stereo_submit(*pairs[0], *frames[0], mbf, mb)
    for i in range(len(frames)):
        if i + 1 < len(frames):
            frame_stereo_submit(*pairs[(i + 1) % 2], *frames[i + 1], mbf, mb)
        (kl, dl), (kr, dr), ur, dp = frame_stereo_wait(*pairs[i % 2])
        assert kl.tobytes() == wantf[i][0].tobytes() and ur.tobytes() == wantf[i][1].tobytes() and dp.tobytes() == wantf[i][2].tobytes()
    # misuse is reported, not ignored
    extract_stereo_submit(*pairs[0], *frames[0])
    with pytest.raises(_lib.CorbError):
        extract_stereo_submit(*pairs[0], *frames[1])
    with pytest.raises(_lib.CorbError):
        frame_stereo_wait(*pairs[0])
    extract_stereo_wait(*pairs[0])
    for a, b in pairs + [ref_pair]:
        a.close(); b.close()


def test_host_transfer_modes_agree():
    """corb_orb_set_host_transfer: mapped reads by the import kernel and copy-engine memcpy nodes give the same stereo pair,
    through the blocking call and through submit / wait, and the mode can be switched on live handles."""
    from corb_slam_b200 import extract_stereo, extract_stereo_submit, extract_stereo_wait
    left, right = stereo_frame(1262)
    exl, exr = ORBextractor(*PARAMS), ORBextractor(*PARAMS)
    (k0l, d0l), (k0r, d0r) = extract_stereo(exl, exr, left, right)
    assert exl.host_transfer() == 0
    exl.set_host_transfer(1); exr.set_host_transfer(1)
    assert exl.host_transfer() == 1
    (k1l, d1l), (k1r, d1r) = extract_stereo(exl, exr, left, right)
    extract_stereo_submit(exl, exr, left, right)
    (k2l, d2l), (k2r, d2r) = extract_stereo_wait(exl, exr)
    for k, d in ((k1l, d1l), (k2l, d2l)):
        assert k.tobytes() == k0l.tobytes() and np.array_equal(d, d0l)
    for k, d in ((k1r, d1r), (k2r, d2r)):
        assert k.tobytes() == k0r.tobytes() and np.array_equal(d, d0r)
    exl.set_host_transfer(0); exr.set_host_transfer(0)
    (k3l, d3l), (k3r, d3r) = extract_stereo(exl, exr, left, right)
    assert k3l.tobytes() == k0l.tobytes() and np.array_equal(d3r, d0r)
