"""The CUDA path against the REFERENCE'S OWN CODE (prebuilt oracle/_ref/libref.so: unmodified ORBextractor.cc, Frame.cc,
ORBmatcher.cc, DBoW2, PnPsolver.cc) - no oracle in between - through the C ABI, on the bench frames and the real ORBvoc.txt
(corbslam_client/Vocabulary, loaded at System.cc:59-68; k = 10, L = 6, 1 082 073 nodes)."""
import numpy as np
import pytest

from oracle import _match_bind as M
from oracle import ref
from corb_slam_b200 import BowFeatures, ORBextractor, ORBmatcher, ORBVocabulary, PnPsolver
from corb_slam_b200.orbextractor import frame_stereo
from corb_slam_b200.synth import KITTI_CAM, pnp_problem, stereo_frame

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref.so was not shipped")]
ORB = (2000, 1.2, 8, 20, 7)


@pytest.mark.parametrize("seed", [1234, 1235])
def test_stereo_frame_equals_the_reference_frame_constructor(seed):
    """corb_frame_stereo (ExtractORB left + right + ComputeStereoMatches in one call) == Frame::Frame(imLeft, imRight, ...)."""
    fx, fy, cx, cy, bf = KITTI_CAM
    L, R = stereo_frame(seed)
    rl, rr = ref.ORBextractor(*ORB), ref.ORBextractor(*ORB)
    F = ref.Frame(rl, rr, L, R, fx, fy, cx, cy, bf)
    gl, gr = ORBextractor(*ORB, device=0), ORBextractor(*ORB, device=0)
    mb = np.float32(bf) / np.float32(fx)
    (kl, dl), (kr, dr), ur, dp = frame_stereo(gl, gr, L, R, bf, mb)
    assert kl.tobytes() == F.keys.tobytes() and kr.tobytes() == F.keys_right.tobytes()
    assert np.array_equal(dl, F.desc) and np.array_equal(dr, F.desc_right)
    assert ur.tobytes() == F.u_right.tobytes() and dp.tobytes() == F.depth.tobytes()
    gl(L, want_pyramid=True)
    for level in range(8):
        assert np.array_equal(gl.mvImagePyramid[level], rl.pyramid(level))
    # getters (ORBextractor.h:63-85)
    assert gl.GetLevels() == 8 and gl.GetScaleFactor() == np.float32(1.2)
    assert gl.GetScaleFactors().tobytes() == rl.scale.tobytes() and gl.GetInverseScaleFactors().tobytes() == rl.inv_scale.tobytes()
    assert gl.GetScaleSigmaSquares().tobytes() == rl.sigma2.tobytes()
    assert gl.GetInverseScaleSigmaSquares().tobytes() == rl.inv_sigma2.tobytes()
    gl.close(); gr.close()


@pytest.fixture(scope="module")
def vocs():
    g = ORBVocabulary(device=0)
    assert g.loadFromTextFile(ref.vocabulary_text())
    r = ref.ORBVocabulary(ref.vocabulary_text(stripped=True))
    yield g, r
    g.close()


@pytest.fixture(scope="module")
def frames():
    ex = ORBextractor(*ORB, device=0)
    base = stereo_frame(1234)[0]
    rng = np.random.default_rng(1)
    again = (np.roll(base, 5, axis=1).astype(np.int16) + rng.normal(0, 3.0, base.shape).round().astype(np.int16)).clip(0, 255).astype(np.uint8)
    out = [tuple(a.copy() for a in ex(i)) for i in (base, again, stereo_frame(1235)[0], stereo_frame(1236)[1])]
    ex.close()
    return out


def test_real_vocabulary_transform_and_score_equal_dbow2(vocs, frames):
    g, r = vocs
    assert (g.k, g.L, g.n_nodes, g.n_words) == (10, 6, 1082073, 971814) and (r.k, r.L, r.n_words) == (10, 6, 971814)
    bows = []
    for _, d in frames:
        a, b = g.transform(d, 4), r.transform(d, 4)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and x.tobytes() == y.tobytes()
        bows.append(b)
    got = g.score_batch(bows[0][:2], [b[:2] for b in bows])
    exp = np.array([r.score(bows[0][:2], b[:2]) for b in bows])
    assert got.tobytes() == exp.tobytes() and exp[1] > 5 * exp[2]


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_search_by_bow_on_the_real_vocabulary_equals_the_reference(vocs, frames, variant):
    g, r = vocs
    fvs = [r.transform(d, 4)[2:] for _, d in frames]
    rng = np.random.default_rng(variant)
    pairs, As, Bs = ((0, 1), (1, 0), (0, 2), (2, 3), (0, 0)), [], []
    for i, j in pairs:
        va, vb = rng.random(len(frames[i][1])) < 0.7, rng.random(len(frames[j][1])) < 0.7
        As.append((frames[i][1], fvs[i], va, frames[i][0]["angle"]))
        Bs.append((frames[j][1], fvs[j], vb, frames[j][0]["angle"]))
    mk = lambda cls, s: cls(s[0], *s[1], valid=s[2], angles=s[3])
    total = 0
    for nn, ori in ((0.7, True), (0.9, False)):
        m = ORBmatcher(nn, ori)
        got = m.SearchByBoWBatch(variant, [mk(BowFeatures, a) for a in As], [mk(BowFeatures, b) for b in Bs])
        for a, b, (gm, gn) in zip(As, Bs, got):
            em, en = ref.search_by_bow(variant, mk(M.Side, a), mk(M.Side, b), nn, ori)
            assert gn == en and np.array_equal(gm, em)
            total += en
        m.close()
    assert total > 2000


def test_pnp_batch_equals_the_reference_solver_on_its_rand_stream():
    sig = ref.ORBextractor(*ORB).sigma2
    gs, rs, seeds = [], [], []
    for c in range(12):
        n = [150, 60, 333, 31, 1000, 97][c % 6]
        p = pnp_problem(100 + c, n=n, outlier_fraction=[0.25, 0.45, 0.1, 0.97][c % 4], pixel_noise=0.5)
        octave = (np.arange(n) % 8).astype(np.int32)
        r = ref.PnPsolver(p["p2d"], octave, p["p3d"], sig, *[float(v) for v in p["K"]])
        r.SetRansacParameters(0.99, 10, 300, 4, 0.5, 5.991)
        g = PnPsolver(p["p2d"], octave, sig, p["K"], p["p3d"], np.ones(n, bool))
        g.SetRansacParameters(0.99, 10, 300, 4, 0.5, 5.991)
        g.set_draws(ref.rand_draws(1000 + c, n, 800))
        gs.append(g); rs.append(r); seeds.append(1000 + c)
    res = PnPsolver.iterate_batch(gs, 5)
    found = 0
    for g, r, seed, (Tcw, no_more, inl, n_inl) in zip(gs, rs, seeds, res):
        rc, rno, rinl, rn, rT = r.iterate(5, seed)
        assert (Tcw is not None) == bool(rc) and no_more == rno and n_inl == rn and g.mnIterations == r.iterations
        assert np.array_equal(inl, rinl)
        if rc:
            assert Tcw.tobytes() == rT.tobytes()
            found += 1
    assert found >= 6
