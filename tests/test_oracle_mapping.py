"""Mapping-thread oracle functions (oracle/mapping_oracle.cpp) against direct numpy restatements.  CPU only."""
import numpy as np

import oracle

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def test_distinctive_descriptors_vs_numpy():
    rng = np.random.default_rng(3)
    sizes = [0, 1, 2, 3, 7, 8, 33, 64, 100, 5, 0, 12]
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = rng.integers(0, 256, (start[-1], 32), dtype=np.uint8)
    desc[start[4]:start[4] + 3] = desc[start[4]]                  # duplicates -> tied medians
    best = oracle.distinctive_descriptors(start, desc)
    for p, n in enumerate(sizes):
        if n == 0:
            assert best[p] == -1
            continue
        d = desc[start[p]:start[p + 1]]
        D = POP[d[:, None, :] ^ d[None, :, :]].sum(axis=2)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert best[p] == int(np.argmin(med))


def test_search_for_triangulation_vs_numpy():
    """orc_search_for_triangulation against a direct numpy loop over the reference's logic on random key frames (descriptor
    clusters so that matches exist), all four flag combinations."""
    rng = np.random.default_rng(11)
    n1, n2, n_nodes = 400, 450, 25
    base = rng.integers(0, 256, (60, 32), dtype=np.uint8)
    def frame(n):
        which = rng.integers(0, len(base), n)
        d = base[which] ^ (rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8))
        k = np.zeros(n, oracle.KP_DTYPE)
        k["x"] = rng.uniform(0, 1241, n); k["y"] = rng.uniform(0, 376, n); k["octave"] = rng.integers(0, 8, n); k["angle"] = rng.uniform(0, 360, n)
        node = which % n_nodes                                   # features of one cluster share a vocabulary node
        order = np.argsort(node, kind="stable")
        ids, start = np.unique(node[order], return_index=True)
        fv = (ids.astype(np.uint32), np.r_[start, n].astype(np.int32), order.astype(np.int32))
        return dict(desc=d, keys=k, has_mp=rng.random(n) < 0.3, uright=np.where(rng.random(n) < 0.5, rng.uniform(0, 1000, n), -1).astype(np.float32), fv=fv)
    kf1, kf2 = frame(n1), frame(n2)
    F12 = rng.normal(0, 1e-3, 9).astype(np.float32); ep = np.array([600.0, 180.0], np.float32)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32); sg = (sf * sf).astype(np.float32)
    for only_stereo, coarse, ori in ((0, 1, 1), (1, 1, 0), (0, 0, 1), (1, 0, 1)):
        nm, match = oracle.search_for_triangulation(kf1, kf2, F12, ep, sf, sg, only_stereo, coarse, ori)
        exp = np.full(n1, -1, np.int32); hist = [[] for _ in range(30)]
        ids1, st1, ft1 = kf1["fv"]; ids2, st2, ft2 = kf2["fv"]
        for a, nid in enumerate(ids1):
            bsel = np.nonzero(ids2 == nid)[0]
            if not len(bsel):
                continue
            b = bsel[0]
            for i1 in ft1[st1[a]:st1[a + 1]]:
                if kf1["has_mp"][i1] or (only_stereo and kf1["uright"][i1] < 0):
                    continue
                best, bi = 50, -1
                for i2 in ft2[st2[b]:st2[b + 1]]:
                    if kf2["has_mp"][i2] or (only_stereo and kf2["uright"][i2] < 0):
                        continue
                    d = int(POP[kf1["desc"][i1] ^ kf2["desc"][i2]].sum())
                    if d > 50 or d > best:
                        continue
                    k1, k2 = kf1["keys"][i1], kf2["keys"][i2]
                    if kf1["uright"][i1] < 0 and kf2["uright"][i2] < 0:
                        ex = np.float32(ep[0] - k2["x"]); ey = np.float32(ep[1] - k2["y"])
                        if np.float32(ex * ex + ey * ey) < np.float32(100 * sf[k2["octave"]]):
                            continue
                    ok = bool(coarse)
                    if not ok:
                        f = np.float32
                        la = f(f(f(k1["x"] * F12[0]) + f(k1["y"] * F12[3])) + F12[6]); lb = f(f(f(k1["x"] * F12[1]) + f(k1["y"] * F12[4])) + F12[7])
                        lc = f(f(f(k1["x"] * F12[2]) + f(k1["y"] * F12[5])) + F12[8])
                        num = f(f(f(la * k2["x"]) + f(lb * k2["y"])) + lc); den = f(f(la * la) + f(lb * lb))
                        ok = den != 0 and float(f(f(num * num) / den)) < 3.84 * float(sg[k2["octave"]])
                    if ok:
                        best, bi = d, i2
                if bi >= 0:
                    exp[i1] = bi
                    rot = np.float32(kf1["keys"]["angle"][i1] - kf2["keys"]["angle"][bi])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360))
                    bn = int(np.round(np.float32(rot * np.float32(1.0 / 30))))
                    hist[0 if bn == 30 else bn].append(i1)
        if ori:
            cnt = [len(h) for h in hist]
            order = sorted([i for i in range(30) if cnt[i] > 0], key=lambda i: (-cnt[i], i))[:3]
            keep = list(order)
            if len(keep) >= 2 and cnt[keep[1]] < np.float32(0.1) * np.float32(cnt[keep[0]]):
                keep = keep[:1]
            elif len(keep) >= 3 and cnt[keep[2]] < np.float32(0.1) * np.float32(cnt[keep[0]]):
                keep = keep[:2]
            for bnum in range(30):
                if bnum not in keep:
                    for i1 in hist[bnum]:
                        exp[i1] = -1
        assert np.array_equal(match, exp), (only_stereo, coarse, ori, int((match != exp).sum()))
        assert nm == int((exp >= 0).sum())
        if coarse:
            assert nm > 20
