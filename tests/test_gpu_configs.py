"""BASELINE.json configs beyond the headline one, as GPU parity tests through the C ABI:
  configs[0]  1241x376 RGB-L, nFeatures = 1000, one frame per call (latency case)
  configs[2]  stereo pair with a smooth NON-uniform disparity field: ORBextractor x2 + Frame::ComputeStereoMatches, then
              SearchByProjection with th = 7 (System::STEREO, src/Tracking.cc:2913-2915) + PoseOptimization on the stereo depths
  configs[3]  1920x1080 RGB-L, ~200 k LiDAR points, nFeatures = 4000: frame construction and the full tracking chain
and the RGB-D depth association (Frame::ComputeStereoFromRGBD, System::TrackRGBD)."""
import numpy as np
import pytest

import oracle
import tracking_data as TD
from oracle import chain as OC
from orb_slam3_rgbl_b200 import frontend as F
from orb_slam3_rgbl_b200 import synthetic as S

pytestmark = pytest.mark.gpu

CAM_D = (1100.0, 1100.0, 960.0, 540.0, 153.0)


def _cmp_kps(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for f in a.dtype.names:
        assert (a[f] == b[f]).all(), f


def test_config_d_1920x1080_200k_points_frame_construction_and_chain():
    W, H, nfeat, T = 1920, 1080, 4000, 4
    seq = S.PlaneSequence(61, T + 1, W=W, H=H, n_azimuth=3125, cam=CAM_D)
    imgs = [seq.image(t) for t in range(T)]; pcs = [seq.cloud(t) for t in range(T)]
    assert pcs[0].shape[1] == 200000
    c = F.Context(W, H, nfeat, max_batch=T, max_points=pcs[0].shape[1])
    try:
        prm = F.make_depth_params(bf=CAM_D[4])
        b = F.RgblBatch(c, imgs, pcs, seq.P, prm, pinned=False)
        b.run_e2e()
        got = [(b.kps[f, :b.n[f]].copy(), b.desc[f, :b.n[f]].copy(), b.depth[f, :b.n[f]].copy(), b.uright[f, :b.n[f]].copy()) for f in range(T)]
        b.track_begin2(F.make_chain_params(seq.pose(0), *CAM_D, th_last=15.0, local_map_frames=2, th_local=3.0))
        g = b.track_end2()
    finally:
        c.close()
    ex = oracle.Extractor(nfeat)
    mask = S.structuring_element("diamond", 5)
    frames = []
    for t in range(T):
        k, d, _ = ex(imgs[t])
        dep, ur, _, _ = oracle.depth_from_pcd(pcs[t], seq.P, W, H, mask, CAM_D[4], k, k)
        frames.append(dict(k=k, d=d, depth=dep, ur=ur))
        _cmp_kps(got[t][0], k)
        assert (got[t][1] == d).all() and (got[t][2] == dep).all() and (got[t][3] == ur).all()
        assert len(k) > 3500 and (dep > 0).sum() > 1000
    rp, rnm, rni, rnl, rni1, _ = OC.oracle_chain2(frames, ex.scale_factors.copy(), seq.pose(0), W, H, CAM_D, K=2)
    # the first tracked frames see identical inputs (see test_gpu_tracking.py for why later ones may not)
    assert g["n_matches"][1] == rnm[1] and g["n_inliers"][1] == rni[1] and g["n_inliers_first"][1] == rni1[1]
    assert np.abs(g["poses"][1] - rp[1]).max() < 2e-4
    assert np.abs(g["n_matches"] - rnm).max() <= 8 and np.abs(g["n_inliers"] - rni).max() <= 12 and np.abs(g["poses"] - rp).max() < 3e-3
    for t in range(T):
        assert abs(g["poses"][t, 4] - seq.pose(t)[4]) < 0.03
    assert g["n_matches"][1:].min() > 400


def test_config_c_stereo_nonuniform_disparity_and_tracking_th7():
    W, H = S.KITTI_W, S.KITTI_H
    left, right = S.stereo_pair(71, W, H)
    el, er = oracle.Extractor(2000), oracle.Extractor(2000)
    kl, dl, _ = el(left); kr, dr, _ = er(right)
    mb = np.float32(S.KITTI_BF) / np.float32(S.KITTI_FX); mbf = np.float32(S.KITTI_BF)
    rd, ru = oracle.stereo_matches(kl, dl, kr, dr, el, er, mb, mbf)
    ex = F.ORBextractor(2000, 1.2, 8, 12, 7, W, H, max_batch=2)
    try:
        (gkl, gdl), (gkr, gdr), gd, gu = F.compute_stereo_matches(ex, (left, right), float(mb), float(mbf))
        _cmp_kps(gkl, kl); _cmp_kps(gkr, kr)
        assert (gd == rd).all() and (gu == ru).all(), f"{(gd != rd).sum()} depths differ"
        ok = rd > 0
        disp = kl["x"][ok] - ru[ok]
        field = S.stereo_disparity_field(W, H)
        truth = field[np.clip(kl["y"][ok].astype(int), 0, H - 1), np.clip((kl["x"][ok] - disp).astype(int), 0, W - 1)]
        assert ok.sum() > 700 and np.median(np.abs(disp - truth)) < 0.6 and disp.max() - disp.min() > 10          # a real disparity range
        # TrackWithMotionModel of a stereo frame against itself displaced: th = 7 (src/Tracking.cc:2913-2915), stereo edges from mvuRight
        sf = el.scale_factors.copy()
        fr = dict(k=kl, d=dl, depth=rd, ur=ru)
        pose0 = np.array([0, 0, 0, 1, 0, 0, 0], np.float32); pose1 = np.array([0, 0, 0, 1, 0.02, -0.01, 0.03], np.float32)
        xw, okp = TD.chain_unproject(fr, pose0)
        ofv = oracle.FrameView(*TD.frame_view_args(fr, sf)); gfv = F.FrameView(*TD.frame_view_args(fr, sf))
        rn, rm = oracle.search_by_projection_last(ofv, pose1, pose0, okp.astype(np.uint8), xw, dl, kl["octave"], kl["angle"], np.ones(len(okp), np.uint8), 7.0)
        n, m = F.ORBmatcher(ex.ctx, 0.9, True).SearchByProjectionLastFrame(gfv, pose1, pose0, okp.astype(np.uint8), xw, dl, kl["octave"], kl["angle"],
                                                                           np.ones(len(okp), np.uint8), 7.0)
        assert n == rn and (m == rm).all() and n > 500
        ix = np.nonzero(rm >= 0)[0]
        obs = np.stack([kl["x"][ix], kl["y"][ix], ru[ix]], 1).astype(np.float32)
        inv_s2 = (np.float32(1) / (sf[kl["octave"][ix]] ** 2).astype(np.float32)).astype(np.float32)
        st = (ru[ix] >= 0).astype(np.uint8)
        rni, rpose, rout = oracle.pose_optimize(pose1, xw[rm[ix]], obs, inv_s2, st, *TD.CAM)
        gni, gpose, gout = F.Optimizer.PoseOptimization(ex.ctx, pose1, xw[rm[ix]], obs, inv_s2, st, *TD.CAM)
        assert gni == rni and (gout == rout).all() and np.abs(gpose - rpose).max() < 2e-5
        assert np.abs(gpose[4:]).max() < 5e-3                     # the frame is matched against itself: the optimum is the identity
    finally:
        ex.ctx.close()


def test_config_a_single_frame_nfeatures_1000():
    W, H = S.KITTI_W, S.KITTI_H
    seq = S.PlaneSequence(81, 3)
    c = F.Context(W, H, 1000, max_batch=1, max_points=seq.cloud(0).shape[1])
    try:
        prm = F.make_depth_params(bf=S.KITTI_BF)
        ex = oracle.Extractor(1000); mask = S.structuring_element("diamond", 5)
        for t in range(2):
            (k, d, dep, ur), = F.frame_rgbl_batch(c, [seq.image(t)], [seq.cloud(t)], seq.P, prm)
            rk, rd, _ = ex(seq.image(t))
            rdep, rur, _, _ = oracle.depth_from_pcd(seq.cloud(t), seq.P, W, H, mask, S.KITTI_BF, rk, rk)
            _cmp_kps(k, rk)
            assert (d == rd).all() and (dep == rdep).all() and (ur == rur).all() and 900 < len(k) <= 1024
    finally:
        c.close()


def test_rgbd_depth_association():
    """Frame::ComputeStereoFromRGBD: the oracle is the reference's own function body (oracle.ref_stereo_from_rgbd) when available."""
    W, H = S.KITTI_W, S.KITTI_H
    rng = np.random.default_rng(5)
    dm = np.where(rng.random((H, W)) < 0.7, rng.uniform(0.5, 80, (H, W)), 0).astype(np.float32)
    n = 1500
    kp = np.zeros(n, oracle.KP_DTYPE); kp["x"] = rng.uniform(0, W - 1, n); kp["y"] = rng.uniform(0, H - 1, n)
    ku = kp.copy(); ku["x"] += rng.normal(0, 0.4, n).astype(np.float32)
    c = F.Context(W, H, 2000, max_batch=1, max_points=1000)
    try:
        d, u = F.compute_stereo_from_rgbd(c, dm, kp, ku, S.KITTI_BF)
    finally:
        c.close()
    v = dm[kp["y"].astype(np.int32), kp["x"].astype(np.int32)]
    ed = np.where(v > 0, v, np.float32(-1)); eu = np.where(v > 0, ku["x"] - np.float32(S.KITTI_BF) / np.where(v > 0, v, 1).astype(np.float32), np.float32(-1)).astype(np.float32)
    assert (d == ed).all() and (u == eu).all() and (d > 0).sum() > 800
    if oracle.ref_tracking() is not None:
        rd, ru = oracle.ref_stereo_from_rgbd(np.stack([kp["x"], kp["y"]], 1), np.stack([ku["x"], ku["y"]], 1), dm, S.KITTI_BF)
        assert (d == rd).all() and (u == ru).all()
