"""ComputeBoW oracle (oracle/bow_oracle.cpp) against an independent numpy/dict restatement of DBoW2's transform
(TemplatedVocabulary.h:1127-1259): descent with first-minimum ties, std::map ordering, addWeight accumulation order,
L1 normalisation, stopped words.  CPU only."""
import numpy as np
import pytest

import oracle
import bow_data as B

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def brute(vocab, desc, levelsup):
    cb, ci, nd, nw, wi, L = (vocab[k] for k in ("child_begin", "child_index", "node_desc", "node_weight", "word_id", "levels"))
    bow, fv = {}, {}
    nid_level = L - levelsup
    for f, q in enumerate(desc):
        node, level, nid = 0, 0, 0
        while cb[node + 1] > cb[node]:
            level += 1
            ch = ci[cb[node]:cb[node + 1]]
            d = POP[nd[ch] ^ q].sum(axis=1)
            node = int(ch[int(np.argmin(d))])          # numpy argmin = first minimum
            if level == nid_level:
                nid = node
        w = float(nw[node])
        if w > 0:
            bow[int(wi[node])] = bow.get(int(wi[node]), 0.0) + w if int(wi[node]) in bow else w
            fv.setdefault(nid, []).append(f)
    words = sorted(bow)
    norm = 0.0
    for k in words:
        norm += abs(bow[k])
    vals = np.array([bow[k] / norm if norm > 0 else bow[k] for k in words], np.float64)
    nodes = sorted(fv)
    start = np.cumsum([0] + [len(fv[k]) for k in nodes]).astype(np.int32)
    feats = np.array([i for k in nodes for i in fv[k]], np.int32)
    return (np.array(words, np.int32), vals), (np.array(nodes, np.int32), start, feats)


@pytest.mark.parametrize("seed,k,levels,ragged,n,levelsup", [(1, 10, 3, False, 700, 2), (2, 10, 3, False, 300, 4), (3, 6, 4, True, 900, 2),
                                                            (4, 10, 2, True, 50, 1), (5, 3, 5, True, 1200, 3), (6, 10, 3, False, 0, 2)])
def test_oracle_equals_brute_force(seed, k, levels, ragged, n, levelsup):
    v = B.make_vocabulary(seed, k, levels, ragged)
    d = B.descriptors_near_words(v, n, seed + 100) if n else np.zeros((0, 32), np.uint8)
    (bw, bv), (fn, fs, ff) = oracle.compute_bow(v, d, levelsup)
    (rw, rv), (rn, rs, rf) = brute(v, d, levelsup)
    assert np.array_equal(bw, rw) and np.array_equal(bv, rv)          # doubles: same additions in the same order -> identical
    assert np.array_equal(fn, rn) and np.array_equal(fs, rs) and np.array_equal(ff, rf)
    if n:
        assert abs(bv.sum() - 1.0) < 1e-12 and (np.diff(bw) > 0).all() and (np.diff(fn) > 0).all()
        stopped = v["node_weight"][v["word_id"] >= 0] == 0
        assert len(ff) <= n and (stopped.any() or len(ff) == n)


# ---- pinned against the reference's own DBoW2 (oracle/_ref/libref_dbow2.so = Thirdparty/DBoW2 compiled unmodified) -------------
@pytest.mark.parametrize("seed,k,levels,ragged,levelsup", [(0, 10, 3, False, 2), (1, 8, 4, True, 3), (2, 10, 4, False, 4), (3, 5, 5, True, 4),
                                                          (4, 10, 3, True, 1), (5, 3, 6, False, 4)])
def test_bow_oracle_equals_the_reference_dbow2(tmp_path, seed, k, levels, ragged, levelsup):
    """The synthetic vocabulary goes through the reference's text loader (ORBvoc.txt format) and its transform(); word ids,
    node ids, feature lists must be identical and the BowVector values bit-identical doubles."""
    ref = oracle.ref_dbow2()
    if ref is None:
        pytest.skip("oracle/_ref/libref_dbow2.so is not built and /root/reference is not available")
    vocab = B.make_vocabulary(seed, k=k, levels=levels, ragged=ragged, shuffle_children=False)
    path = tmp_path / "voc.txt"
    oracle.write_vocabulary_text(vocab, path, k)
    h = ref.ref_voc_load_text(str(path).encode())
    assert h
    try:
        assert ref.ref_voc_size(h) == vocab["n_words"]
        for n in (0, 1, 37, 1200):
            desc = B.descriptors_near_words(vocab, n, 50 + seed) if n else np.zeros((0, 32), np.uint8)
            (w, v), (fn, fs, ff) = oracle.compute_bow(vocab, desc, levelsup)
            (rw, rv), (rfn, rfs, rff) = oracle.ref_compute_bow(h, desc, levelsup)
            assert (w == rw).all() and len(w) == len(rw)
            assert (v.view(np.uint64) == rv.view(np.uint64)).all()
            assert (fn == rfn).all() and (fs == rfs).all() and (ff == rff).all()
    finally:
        ref.ref_voc_free(h)
