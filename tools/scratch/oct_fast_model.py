"""Python model of the closed-form quadtree path of k_octtree (design check, CPU only).

ref_distribute  : list-based restatement (same as oracle/orb_oracle.cpp distribute)
fast_distribute : histogram / closed-form formulation the CUDA kernel implements; returns None when it would bail
                  to the generic path.
Run: python tools/scratch/oct_fast_model.py
"""
import math
import sys
import os

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))

DH = 5


def f32(x):
    return np.float32(x)


def root_of(x, hX):
    return int(f32(x) / hX)


def ref_distribute(keys, width, height, N):
    nIni = int(round(float(f32(width) / f32(height))))
    # std::round(float): half away from zero
    v = float(f32(width) / f32(height))
    nIni = int(math.floor(v + 0.5))
    hX = f32(width) / f32(nIni)
    seq = [0]

    class Node:
        __slots__ = ("x0", "y0", "x1", "y1", "keys", "no_more", "seq")

    nodes = []
    ini = []
    for i in range(nIni):
        n = Node()
        n.x0, n.y0, n.x1, n.y1 = int(hX * f32(i)), 0, int(hX * f32(i + 1)), height
        n.keys, n.no_more, n.seq = [], False, seq[0]
        seq[0] += 1
        nodes.append(n)
        ini.append(n)
    for k in keys:
        ini[root_of(k[0], hX)].keys.append(k)
    nn = []
    for n in nodes:
        if len(n.keys) == 1:
            n.no_more = True
            nn.append(n)
        elif len(n.keys) > 1:
            nn.append(n)
    nodes = nn

    def divide(n):
        halfX = (n.x1 - n.x0 + 1) // 2
        halfY = (n.y1 - n.y0 + 1) // 2
        xm, ym = n.x0 + halfX, n.y0 + halfY
        cs = []
        for (a, b, c, d) in ((n.x0, n.y0, xm, ym), (xm, n.y0, n.x1, ym), (n.x0, ym, xm, n.y1), (xm, ym, n.x1, n.y1)):
            m = Node()
            m.x0, m.y0, m.x1, m.y1, m.keys, m.no_more, m.seq = a, b, c, d, [], False, 0
            cs.append(m)
        for k in n.keys:
            q = (0 if k[0] < xm else 1) + (0 if k[1] < ym else 2)
            cs[q].keys.append(k)
        for m in cs:
            if len(m.keys) == 1:
                m.no_more = True
        return cs

    finish = False
    while not finish:
        prev = len(nodes)
        nexp = 0
        expandable = []
        new_front = []
        rest = []
        for n in nodes:
            if n.no_more:
                rest.append(n)
                continue
            for c in divide(n):
                if not c.keys:
                    continue
                c.seq = seq[0]
                seq[0] += 1
                new_front.insert(0, c)
                if len(c.keys) > 1:
                    nexp += 1
                    expandable.append(c)
        nodes = new_front + rest
        if len(nodes) >= N or len(nodes) == prev:
            finish = True
        elif len(nodes) + 3 * nexp > N:
            while not finish:
                prev = len(nodes)
                pv = sorted(expandable, key=lambda n: (len(n.keys), n.seq))
                expandable = []
                for n in reversed(pv):
                    for c in divide(n):
                        if not c.keys:
                            continue
                        c.seq = seq[0]
                        seq[0] += 1
                        nodes.insert(0, c)
                        if len(c.keys) > 1:
                            expandable.append(c)
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if k[2] > best[2]:
                best = k
        out.append(tuple(best))
    return out


def fast_distribute(keys, width, height, N, dh=DH):
    v = float(f32(width) / f32(height))
    nIni = int(math.floor(v + 0.5))
    hX = f32(width) / f32(nIni)
    M = len(keys)
    if M == 0:
        return None
    HC = nIni << (2 * dh)
    hist = np.zeros(HC + 1, np.int64)
    code = np.zeros(M, np.int64)
    for i, k in enumerate(keys):
        r = root_of(k[0], hX)
        x0, y0, x1, y1 = int(hX * f32(r)), 0, int(hX * f32(r + 1)), height
        c = r
        for d in range(dh):
            xm = x0 + (x1 - x0 + 1) // 2
            ym = y0 + (y1 - y0 + 1) // 2
            q = (0 if k[0] < xm else 1) + (0 if k[1] < ym else 2)
            if q & 1: x0 = xm
            else: x1 = xm
            if q & 2: y0 = ym
            else: y1 = ym
            c = c * 4 + q
        code[i] = c
        hist[c] += 1
    P = np.concatenate([[0], np.cumsum(hist[:HC])])

    def cnt(d, c):
        sh = 2 * (dh - d)
        return int(P[(c + 1) << sh] - P[c << sh])

    S = [sum(1 for c in range(nIni << (2 * d)) if cnt(d, c) > 0) for d in range(dh + 1)]
    E = [sum(1 for c in range(nIni << (2 * d)) if cnt(d, c) > 1) for d in range(dh + 1)]
    dstar, mode = -1, None
    sp = S[0]
    for d in range(1, dh + 1):
        s = S[d]
        if s >= N or s == sp:
            dstar, mode = d, "finish"
            break
        if s + 3 * E[d] > N:
            dstar, mode = d, "final"
            break
        sp = s
    if dstar < 0:
        return None

    def tau(d, i):
        r = i >> (2 * d)
        low = i & ((1 << (2 * d)) - 1)
        low ^= 0x33333333 & ((1 << (2 * d)) - 1)
        if d & 1:
            r = nIni - 1 - r
        return (r << (2 * d)) | low

    # node list: (depth, cell, cnt, fresh)
    nodes = []
    for j in range(dstar + 1):
        d = dstar - j
        for i in range(nIni << (2 * d)):
            c = tau(d, i)
            n = cnt(d, c)
            par_ok = d == 0 or cnt(d - 1, c >> 2) > 1
            if j == 0:
                mem = n > 0 and par_ok
            else:
                mem = n == 1 and par_ok
            if mem:
                nodes.append((d, c, n, j == 0))
    assert len(nodes) == S[dstar], (len(nodes), S[dstar])
    if mode == "final":
        finish = False
        while not finish:
            prev = len(nodes)
            cand = [(i, nd) for i, nd in enumerate(nodes) if nd[3] and nd[2] > 1]
            order = sorted(cand, key=lambda t: (-t[1][2], t[0]))
            s = len(nodes)
            created = []  # creation order
            processed = set()
            for pos, nd in order:
                d, c, n, _ = nd
                if d + 1 > dh:
                    return None
                for q in range(4):
                    cn = cnt(d + 1, c * 4 + q)
                    if cn > 0:
                        created.append((d + 1, c * 4 + q, cn, True))
                        s += 1
                processed.add(pos)
                s -= 1
                if s >= N:
                    break
            rest = [(nd[0], nd[1], nd[2], False) for i, nd in enumerate(nodes) if i not in processed]
            nodes = list(reversed(created)) + rest
            if len(nodes) >= N or len(nodes) == prev:
                finish = True
    # owner + best
    owner = {}
    for pos, nd in enumerate(nodes):
        owner[(nd[0], nd[1])] = pos
    best = [None] * len(nodes)
    for i, k in enumerate(keys):
        o = None
        for d in range(dh + 1):
            o = owner.get((d, int(code[i]) >> (2 * (dh - d))))
            if o is not None:
                break
        assert o is not None
        if best[o] is None or k[2] > best[o][2]:
            best[o] = k
    return [tuple(b) for b in best]


def main():
    rng = np.random.default_rng(0)
    n_fast = n_bail = 0
    # 1. random key sets (distinct pixels), various densities / aspect ratios / quotas
    for trial in range(300):
        W = int(rng.integers(60, 1300))
        H = int(rng.integers(40, 400))
        if int(math.floor(float(f32(W) / f32(H)) + 0.5)) < 1:
            continue
        M = int(rng.integers(1, 6000))
        N = int(rng.integers(5, 1200))
        M = min(M, W * H // 4)
        idx = rng.choice(W * H, M, replace=False)
        clustered = trial % 3 == 0
        if clustered:
            cx, cy = rng.integers(0, W), rng.integers(0, H)
            xs = np.clip((rng.normal(cx, W / 8, M)).astype(int), 0, W - 1)
            ys = np.clip((rng.normal(cy, H / 8, M)).astype(int), 0, H - 1)
            pts = sorted(set(zip(xs.tolist(), ys.tolist())))
            rng.shuffle(pts)
            keys = [(x, y, int(rng.integers(7, 100))) for x, y in pts]
        else:
            keys = [(int(i % W), int(i // W), int(rng.integers(7, 100))) for i in idx]
        a = ref_distribute(keys, W, H, N)
        b = fast_distribute(keys, W, H, N)
        if b is None:
            n_bail += 1
            continue
        n_fast += 1
        assert a == b, (trial, W, H, M, N, len(a), len(b))
    print("random: fast %d, bail %d, all equal" % (n_fast, n_bail))
    # 2. oracle candidates of synthetic frames, per level, checked against the oracle's own output for level 0
    import oracle
    from corb_slam_b200.synth import stereo_frame, frame_seed
    ex = oracle.OrbExtractor(2000, 1.2, 8, 20, 7)
    for fi in range(4):
        l, _ = stereo_frame(frame_seed(fi))
        kps, _ = ex(l)
        off = 0
        for lev in range(8):
            w, h = ex.level_size(lev, l.shape[1], l.shape[0])
            cand = [tuple(int(v) for v in row) for row in ex.candidates(lev)]
            N = int(ex.quota[lev])
            a = ref_distribute(cand, w - 32, h - 32, N)
            b = fast_distribute(cand, w - 32, h - 32, N)
            cnt = ex.level_count(lev)
            assert len(a) == cnt, (lev, len(a), cnt)
            if lev == 0:
                o = [(int(k["x"]) - 16, int(k["y"]) - 16, int(k["response"])) for k in kps[off:off + cnt]]
                assert o == a, "python list model differs from the oracle"
            off += cnt
            print("frame %d level %d: M=%d N=%d -> %d nodes, fast path %s" % (fi, lev, len(cand), N, len(a),
                  "bails" if b is None else ("equal" if a == b else "DIFFERS")))
            assert b is None or a == b


if __name__ == "__main__":
    main()
