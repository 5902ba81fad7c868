"""Mapping-thread kernels on the device against the oracle: MapPoint::ComputeDistinctiveDescriptors (batch) and
ORBmatcher::SearchForTriangulation between two extracted key frames (feature vectors from ComputeBoW)."""
import numpy as np
import pytest

import oracle
import bow_data as B
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=2, max_points=140000)
    yield c
    c.close()


def test_distinctive_descriptors(ctx):
    rng = np.random.default_rng(5)
    sizes = [0, 1, 2, 3, 7, 8, 31, 32, 33, 64, 100, 257, 5, 0, 12] + list(rng.integers(1, 40, 3000))
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    desc = np.repeat(base, sizes, axis=0)
    flips = rng.integers(0, 256, desc.shape, dtype=np.uint8) & rng.integers(0, 256, desc.shape, dtype=np.uint8) & rng.integers(0, 256, desc.shape, dtype=np.uint8)
    desc ^= flips                                                   # observations of one point: noisy copies (many tied medians)
    got = F.distinctive_descriptors(ctx, start, desc)
    ref = oracle.distinctive_descriptors(start, desc)
    assert np.array_equal(got, ref)
    assert np.array_equal(F.distinctive_descriptors(ctx, np.zeros(1, np.int32), np.zeros((0, 32), np.uint8)), np.zeros(0, np.int32))


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float32)


@pytest.mark.parametrize("only_stereo,coarse,check_ori,ep_in_image", [(False, False, True, False), (True, False, True, False), (False, True, False, False),
                                                                      (False, False, False, False), (False, True, True, True)])
def test_search_for_triangulation(ctx, only_stereo, coarse, check_ori, ep_in_image):
    seq = S.PlaneSequence(61, 4)
    imgs = [seq.image(0), seq.image(2)]; pcs = [seq.cloud(0), seq.cloud(2)]
    outs = F.frame_rgbl_batch(ctx, imgs, pcs, seq.P, F.make_depth_params(bf=S.KITTI_BF))
    (k1, d1, dep1, ur1), (k2, d2, dep2, ur2) = outs
    rng = np.random.default_rng(7)
    v = B.make_vocabulary(21, 10, 4)
    leaves = np.nonzero(v["word_id"] >= 0)[0]
    v["node_desc"][leaves[:len(d1)]] = d1[rng.permutation(len(d1))][:len(leaves)]
    fv1 = oracle.compute_bow(v, d1)[1]; fv2 = oracle.compute_bow(v, d2)[1]
    kf1 = dict(desc=d1, keys=k1, has_mp=rng.random(len(d1)) < 0.3, uright=ur1, fv=fv1)
    kf2 = dict(desc=d2, keys=k2, has_mp=rng.random(len(d2)) < 0.3, uright=ur2, fv=fv2)
    # per-pair constants as the shim computes them: T12 = T1w * Tw2, F12 = K^-T [t12]x R12 K^-1, epipole of camera 1 in image 2
    T1, T2 = seq.pose(0).astype(np.float64), seq.pose(2).astype(np.float64)
    R1, R2 = _rot(T1[:4]).astype(np.float64), _rot(T2[:4]).astype(np.float64)
    R12 = R1 @ R2.T; t12 = T1[4:] - R12 @ T2[4:]
    K = np.array([[S.KITTI_FX, 0, S.KITTI_CX], [0, S.KITTI_FY, S.KITTI_CY], [0, 0, 1]])
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = (np.linalg.inv(K).T @ tx @ R12 @ np.linalg.inv(K)).astype(np.float32)
    Cw = -R1.T @ T1[4:]; C2 = R2 @ Cw + T2[4:]
    with np.errstate(all="ignore"):      # the synthetic camera moves sideways: the epipole is at infinity (the reference divides by zero too)
        ep = np.array([S.KITTI_FX * C2[0] / C2[2] + S.KITTI_CX, S.KITTI_FY * C2[1] / C2[2] + S.KITTI_CY], np.float32)
    if ep_in_image:                       # exercise the "too close to the epipole" rejection (mono-mono pairs only)
        ep = np.array([620.0, 188.0], np.float32)
        kf1["uright"] = np.full(len(d1), -1, np.float32); kf2["uright"] = np.full(len(d2), -1, np.float32)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32); sg = (sf * sf).astype(np.float32)
    rnm, rmatch = oracle.search_for_triangulation(kf1, kf2, F12, ep, sf, sg, only_stereo, coarse, check_ori)
    gnm, gmatch = F.search_for_triangulation(ctx, kf1, kf2, F12, ep, sf, sg, only_stereo, coarse, check_ori)
    assert gnm == rnm and np.array_equal(gmatch, rmatch)
    if coarse:
        assert rnm > 50
    m = gmatch >= 0
    assert not kf1["has_mp"][m].any() and not kf2["has_mp"][gmatch[m]].any()
    if only_stereo:
        assert (ur1[m] >= 0).all() and (ur2[gmatch[m]] >= 0).all()
