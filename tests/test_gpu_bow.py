"""Frame::ComputeBoW on the device (rgbl_vocabulary_create / rgbl_compute_bow / rgbl_resident_compute_bow) against the oracle:
word ids, node ids, feature lists identical; BowVector values bit-identical doubles; and the result feeds SearchByBoW."""
import numpy as np
import pytest

import oracle
import bow_data as B
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=2)
    yield c
    c.close()


def _same(got, ref):
    (gw, gv), (gn, gs, gf) = got
    (rw, rv), (rn, rs, rf) = ref
    assert np.array_equal(gw, rw), "word ids differ"
    assert np.array_equal(gv.view(np.uint64), rv.view(np.uint64)), f"BowVector values differ by up to {np.abs(gv - rv).max()}"
    assert np.array_equal(gn, rn) and np.array_equal(gs, rs) and np.array_equal(gf, rf), "FeatureVector differs"


@pytest.mark.parametrize("seed,k,levels,ragged,n,levelsup", [(1, 10, 3, False, 700, 2), (2, 10, 4, False, 2000, 4), (3, 6, 4, True, 900, 2),
                                                            (4, 10, 2, True, 50, 1), (5, 3, 5, True, 1200, 3), (7, 40, 2, False, 333, 1),
                                                            (8, 10, 3, False, 5000, 2), (9, 10, 3, False, 1, 2)])
def test_transform_matches_oracle(ctx, seed, k, levels, ragged, n, levelsup):
    v = B.make_vocabulary(seed, k, levels, ragged)
    d = B.descriptors_near_words(v, n, seed + 100)
    voc = F.ORBVocabulary(ctx, v["child_begin"], v["child_index"], v["node_desc"], v["node_weight"], v["word_id"], v["levels"])
    try:
        _same(voc.transform(d, levelsup), oracle.compute_bow(v, d, levelsup))
    finally:
        voc.close()


def test_edge_cases_and_errors(ctx):
    v = B.make_vocabulary(11, 10, 3)
    voc = F.ORBVocabulary(ctx, v["child_begin"], v["child_index"], v["node_desc"], v["node_weight"], v["word_id"], v["levels"])
    try:
        (bw, bv), (fn, fs, ff) = voc.transform(np.zeros((0, 32), np.uint8))
        assert len(bw) == 0 and len(fn) == 0 and list(fs) == [0]
        allstop = dict(v); allstop["node_weight"] = np.zeros_like(v["node_weight"])
        voc0 = F.ORBVocabulary(ctx, allstop["child_begin"], allstop["child_index"], allstop["node_desc"], allstop["node_weight"], allstop["word_id"], 3)
        try:
            (bw, bv), (fn, fs, ff) = voc0.transform(B.descriptors_near_words(v, 100, 1))
            assert len(bw) == 0 and len(fn) == 0 and len(ff) == 0          # every word stopped
        finally:
            voc0.close()
        with pytest.raises(Exception):
            F.ORBVocabulary(ctx, v["child_begin"], v["child_index"], v["node_desc"], v["node_weight"], v["word_id"], 3, weighting=3)   # BINARY
        with pytest.raises(Exception):
            F.ORBVocabulary(ctx, v["child_begin"], v["child_index"], v["node_desc"], v["node_weight"], v["word_id"], 3, scoring=1)     # L2
        bad = v["child_index"].copy(); bad[0] = len(v["word_id"])
        with pytest.raises(Exception):
            F.ORBVocabulary(ctx, v["child_begin"], bad, v["node_desc"], v["node_weight"], v["word_id"], 3)
        with pytest.raises(Exception):
            voc.transform(np.zeros((9000, 32), np.uint8))               # beyond the one-CTA sort capacity: reported, not truncated
    finally:
        voc.close()


def test_resident_descriptors_and_search_by_bow(ctx):
    """ComputeBoW on two extracted frames straight from HBM, then SearchByBoW driven by those feature vectors = the oracle pair."""
    seq = S.PlaneSequence(31, 3)
    imgs = [seq.image(0), seq.image(1)]
    ex = F.ORBextractor(2000, 1.2, 8, 12, 7, S.KITTI_W, S.KITTI_H, ctx=ctx)
    (k0, d0), (k1, d1) = ex.extract_batch(imgs)
    v = B.make_vocabulary(12, 10, 4)
    # vocabulary words made from the frames' own descriptors so that corresponding features share nodes
    rng = np.random.default_rng(0)
    leaves = np.nonzero(v["word_id"] >= 0)[0]
    v["node_desc"][leaves[:len(d0)]] = d0[rng.permutation(len(d0))][:len(leaves)]
    voc = F.ORBVocabulary(ctx, v["child_begin"], v["child_index"], v["node_desc"], v["node_weight"], v["word_id"], v["levels"])
    try:
        got0, got1 = voc.transform_resident(0), voc.transform_resident(1)
        ref0, ref1 = oracle.compute_bow(v, d0), oracle.compute_bow(v, d1)
        _same(got0, ref0); _same(got1, ref1)
        _same(voc.transform(d1), ref1)
        m = F.ORBmatcher(ctx, 0.7, True)
        valid = np.ones(len(d0), np.uint8)
        nm, match = m.SearchByBoW(d0, k0["angle"], valid, got0[1], d1, k1["angle"], got1[1])
        rnm, rmatch = oracle.search_by_bow(d0, k0["angle"], valid, ref0[1], d1, k1["angle"], ref1[1], 0.7, True)
        assert nm == rnm and np.array_equal(match, rmatch)
    finally:
        voc.close()
