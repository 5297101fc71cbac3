"""GPU parity of Hamming / SearchByBoW x3 / DBoW2 transform / L1 score against the CPU oracle (bit-exact)."""
import numpy as np
import pytest

import oracle
from oracle import _match_bind as M
from corb_slam_b200 import BowFeatures, ORBmatcher, ORBVocabulary

pytestmark = pytest.mark.gpu


def _voc(k=10, L=3, seed=0):
    oracle.lib()
    spec = M.random_vocabulary(k, L, seed)
    return M.Vocabulary.from_arrays(*spec), ORBVocabulary.from_arrays(*spec)


def _features(rng, n, base=None, flip_bits=12):
    """n random descriptors; if `base` is given, noisy copies of it (a second view of the same scene)."""
    if base is None:
        return rng.integers(0, 256, (n, 32), dtype=np.uint8)
    d = base[rng.integers(0, len(base), n)].copy()
    for i in range(n):
        for b in rng.integers(0, 256, rng.integers(0, flip_bits)):
            d[i, b >> 3] ^= 1 << (b & 7)
    return d


def _sides(rng, ovoc, nA, nB, levelsup, p_valid=0.7):
    dA = _features(rng, nA)
    dB = _features(rng, nB, base=dA)
    bA = ovoc.transform(dA, levelsup)
    bB = ovoc.transform(dB, levelsup)
    angA = rng.uniform(0, 360, nA).astype(np.float32)
    angB = ((angA[rng.integers(0, nA, nB)] + rng.normal(0, 20, nB)) % 360).astype(np.float32)
    vA = (rng.random(nA) < p_valid).astype(np.uint8)
    vB = (rng.random(nB) < p_valid).astype(np.uint8)
    return (dA, bA, angA, vA), (dB, bB, angB, vB)


def _mk(cls, s):
    d, b, ang, v = s
    return cls(d, b[2], b[3], b[4], valid=v, angles=ang)


def test_hamming_pairs_bit_exact():
    rng = np.random.default_rng(0)
    A = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    B[:50] = A[:50]
    pairs = np.stack([rng.integers(0, 300, 5000), rng.integers(0, 200, 5000)], 1).astype(np.int32)
    m = ORBmatcher()
    got = m.DescriptorDistancePairs(A, B, pairs)
    exp = np.array([M.hamming256(A[i], B[j]) for i, j in pairs], np.int32)
    np.testing.assert_array_equal(got, exp)
    assert ORBmatcher.DescriptorDistance(A[3], B[7]) == M.hamming256(A[3], B[7])
    assert M.hamming256(A[0], B[0]) == 0 and M.hamming256(A[1], 255 - A[1]) == 256
    m.close()


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("check_ori", [True, False])
@pytest.mark.parametrize("ratio", [0.6, 0.75, 0.9])
def test_search_by_bow_bit_exact(variant, check_ori, ratio):
    ovoc, _ = _voc(10, 3, 1)
    rng = np.random.default_rng(100 + variant)
    m = ORBmatcher(ratio, check_ori)
    for trial in range(4):
        sA, sB = _sides(rng, ovoc, 1500 + 100 * trial, 1800, levelsup=[1, 2, 2, 3][trial])
        exp, en = M.search_by_bow(variant, _mk(M.Side, sA), _mk(M.Side, sB), ratio, check_ori)
        got, gn = m._batch(variant, [_mk(BowFeatures, sA)], [_mk(BowFeatures, sB)])[0]
        assert gn == en
        np.testing.assert_array_equal(got, exp)
        assert en > 10, "test data should produce matches"
    m.close()


def test_search_by_bow_batch_and_edge_cases():
    ovoc, _ = _voc(10, 3, 2)
    rng = np.random.default_rng(5)
    m = ORBmatcher(0.75, True)
    sides = [_sides(rng, ovoc, n, n + 37, 2) for n in (5, 64, 700, 2000)]
    # an empty feature vector on one side, and identical descriptors (ties: best == second -> ratio test fails)
    dA = np.tile(rng.integers(0, 256, (1, 32), dtype=np.uint8), (40, 1))
    b = ovoc.transform(dA, 2)
    tie = ((dA, b, np.zeros(40, np.float32), np.ones(40, np.uint8)),) * 2
    sides.append(tie)
    empty = (np.zeros((0, 32), np.uint8), ovoc.transform(np.zeros((0, 32), np.uint8), 2), np.zeros(0, np.float32), np.zeros(0, np.uint8))
    sides.append((sides[1][0], empty))
    for variant in (0, 2):
        got = m.SearchByBoWBatch(variant, [_mk(BowFeatures, a) for a, _ in sides], [_mk(BowFeatures, b_) for _, b_ in sides])
        for (a, b_), (gm, gn) in zip(sides, got):
            em, en = M.search_by_bow(variant, _mk(M.Side, a), _mk(M.Side, b_), 0.75, True)
            assert gn == en
            np.testing.assert_array_equal(gm, em)
    m.close()


@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 1), (10, 4, 2), (7, 5, 4), (16, 2, 1), (10, 3, 5)])
def test_vocabulary_transform_bit_exact(k, L, levelsup):
    ovoc, gvoc = _voc(k, L, 3)
    assert (gvoc.k, gvoc.L, gvoc.n_nodes, gvoc.n_words) == (ovoc.k, ovoc.L, ovoc.n_nodes, ovoc.n_words)
    rng = np.random.default_rng(9)
    d = _features(rng, 2000)
    ew, ewt, en = ovoc.transform_features(d, levelsup)
    gw, gwt, gn = gvoc.transform_features(d, levelsup)
    np.testing.assert_array_equal(gw, ew)
    np.testing.assert_array_equal(gwt.view(np.uint64), ewt.view(np.uint64))
    np.testing.assert_array_equal(gn, en)
    eb = ovoc.transform(d, levelsup)
    gb = gvoc.transform(d, levelsup)
    for g, e in zip(gb, eb):
        assert g.dtype == e.dtype
        np.testing.assert_array_equal(g.view(np.uint8), e.view(np.uint8))
    gvoc.close()


def test_unbalanced_vocabulary_and_text_loader(tmp_path):
    # leaves at different depths, nodes with 2..10 children (like the real ORBvoc: leaf depth 4..6, 2..10 children)
    rng = np.random.default_rng(4)
    parent, leaf = [], []
    frontier = [(0, 0)]
    nid = 0
    while frontier:
        p, depth = frontier.pop(0)
        for _ in range(int(rng.integers(2, 11))):
            nid += 1
            is_leaf = depth + 1 >= 4 or (depth + 1 >= 2 and rng.random() < 0.3)
            parent.append(p); leaf.append(int(is_leaf))
            if not is_leaf:
                frontier.append((nid, depth + 1))
    n = nid
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    weight = np.where(np.array(leaf) > 0, np.round(rng.uniform(0, 9, n), 5), 0.0)
    path = tmp_path / "voc.txt"
    with open(path, "w") as f:
        f.write("10 4  0 0\n")
        for i in range(n):
            f.write("%d %d %s %s\n" % (parent[i], leaf[i], " ".join(str(int(x)) for x in desc[i]), repr(float(weight[i]))))
    ovoc = M.Vocabulary.load_text(str(path))
    gvoc = ORBVocabulary()
    assert gvoc.loadFromTextFile(str(path))
    assert (gvoc.n_nodes, gvoc.n_words) == (ovoc.n_nodes, ovoc.n_words) == (n + 1, sum(leaf))
    d = _features(rng, 1500)
    for levelsup in (1, 2, 3):
        for g, e in zip(gvoc.transform(d, levelsup), ovoc.transform(d, levelsup)):
            np.testing.assert_array_equal(g.view(np.uint8), e.view(np.uint8))
    bad = tmp_path / "bad.txt"
    bad.write_text("hello world\n")
    assert not ORBVocabulary().loadFromTextFile(str(bad))
    gvoc.close()


def test_l1_score_batch_bit_exact():
    ovoc, gvoc = _voc(10, 3, 6)
    rng = np.random.default_rng(12)
    base = _features(rng, 2000)
    q = ovoc.transform(base, 2)[:2]
    cands = []
    for i in range(64):
        d = _features(rng, int(rng.integers(1, 2000)), base=base if i % 2 else None)
        cands.append(ovoc.transform(d, 2)[:2])
    cands.append((np.zeros(0, np.uint32), np.zeros(0, np.float64)))
    cands.append(q)
    got = gvoc.score_batch(q, cands)
    exp = np.array([M.bow_score_l1(q[0], q[1], c[0], c[1]) for c in cands])
    np.testing.assert_array_equal(got.view(np.uint64), exp.view(np.uint64))
    assert exp[-1] == pytest.approx(1.0) and exp[-2] == 0.0
    assert gvoc.score(q, cands[1]) == exp[1]
    gvoc.close()


def test_device_resident_bow_records_equal_the_host_path():
    """corb_frame_bow / corb_bow_store_*: BowVector + FeatureVector built on the device from the extractor's resident
    descriptors equal corb_voc_transform (and therefore DBoW2) bit for bit; SearchByBoW and the L1 score between records
    equal the host-array entry points."""
    from corb_slam_b200 import BowRecord, ORBextractor
    from corb_slam_b200.synth import stereo_frame
    ovoc, gvoc = _voc(10, 4, 5)
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    rng = np.random.default_rng(2)
    base = stereo_frame(1234)[0]
    imgs = [base, (np.roll(base, 4, axis=1).astype(np.int16) + rng.normal(0, 3, base.shape).round().astype(np.int16)).clip(0, 255).astype(np.uint8),
            stereo_frame(1235)[0]]
    recs, host = [], []
    for img in imgs:
        k, d = ex(img)
        k, d = k.copy(), d.copy()
        r = BowRecord(2048).from_extractor(ex, gvoc, 2)
        got, exp = r.download(), gvoc.transform(d, 2)
        for a, b in zip(got, exp):
            assert a.dtype == b.dtype and a.tobytes() == b.tobytes()
        for a, b in zip(got, ovoc.transform(d, 2)):
            assert a.tobytes() == b.tobytes()
        recs.append(r); host.append((k, d, exp))
    # L1 scores between records
    sc = BowRecord.score(gvoc, recs[0], recs)
    exp = gvoc.score_batch(host[0][2][:2], [h[2][:2] for h in host])
    assert sc.tobytes() == exp.tobytes() and sc[0] == pytest.approx(1.0)
    # SearchByBoW between records vs host arrays, all variants
    valid = [(rng.random(len(h[0])) < 0.7).astype(np.uint8) for h in host]
    for variant in (0, 1, 2):
        m = ORBmatcher(0.75, True)
        pairs = [(0, 1), (1, 0), (0, 2)]
        got = m.SearchByBoWRecords(variant, [recs[i] for i, _ in pairs], [valid[i] for i, _ in pairs], [recs[j] for _, j in pairs],
                                   [valid[j] for _, j in pairs])
        A = [BowFeatures(host[i][1], *host[i][2][2:], valid=valid[i], angles=host[i][0]["angle"]) for i, _ in pairs]
        B = [BowFeatures(host[j][1], *host[j][2][2:], valid=valid[j], angles=host[j][0]["angle"]) for _, j in pairs]
        exp = m.SearchByBoWBatch(variant, A, B)
        for (gm, gn), (em, en) in zip(got, exp):
            assert gn == en and np.array_equal(gm, em)
        assert exp[0][1] > 100
        m.close()
    # an empty record and a record filled from plain device arrays
    import torch
    e = BowRecord(64).from_device(gvoc, None, 0, 2)
    assert all(len(a) == 0 for a in e.download()[:3])
    d = torch.from_numpy(host[2][1][:50].copy()).cuda()
    r = BowRecord(64).from_device(gvoc, d.data_ptr(), 50, 2)
    for a, b in zip(r.download(), gvoc.transform(host[2][1][:50], 2)):
        assert a.tobytes() == b.tobytes()
    for r_ in recs:
        r_.close()
    ex.close(); gvoc.close()
