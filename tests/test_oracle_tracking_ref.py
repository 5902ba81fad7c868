"""Pins the matcher / pose / stereo oracles to the REFERENCE's own code (oracle/_ref/libref_tracking.so = src/ORBmatcher.cc,
Optimizer::PoseOptimization over the reference's g2o, Frame::isInFrustum / GetFeaturesInArea / ComputeStereoMatches /
ComputeStereoFromRGBD / UnprojectStereo, MapPoint::ComputeDistinctiveDescriptors, compiled unmodified over stand-in Eigen / Sophus /
OpenCV headers; see oracle/ref_tracking_driver.cpp).  Both sides receive the same arrays through the same Python wrappers
(`with oracle.reference_tracking():` switches the library).  Integer results must be identical; float32 results bit-identical;
the float32 pose returned by the FP64 Levenberg-Marquardt bit-identical too."""
import numpy as np
import pytest

import oracle
import tracking_data as TD
from orb_slam3_rgbl_b200 import synthetic as S

pytestmark = pytest.mark.skipif(oracle.ref_tracking() is None, reason="oracle/_ref/libref_tracking.so not built (needs /root/reference)")


def both(fn, *a, **k):
    got = fn(*a, **k)
    with oracle.reference_tracking():
        ref = fn(*a, **k)
    return got, ref


@pytest.fixture(scope="module")
def seq_frames():
    seq = S.PlaneSequence(5, 3)
    frames, sf = TD.extract_frames(seq, [0, 1, 2])
    return seq, frames, sf


def rot_pose(rng, t_scale=0.3, r_scale=0.02):
    w = rng.normal(0, r_scale, 3); th = np.linalg.norm(w)
    q = np.r_[w / th * np.sin(th / 2), np.cos(th / 2)] if th > 0 else np.array([0, 0, 0, 1.0])
    return np.r_[q, rng.normal(0, t_scale, 3)].astype(np.float32)


def test_descriptor_distance():
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        got, ref = both(oracle.descriptor_distance, a, b)
        assert got == ref


def test_features_in_area(seq_frames):
    seq, frames, sf = seq_frames
    fv = oracle.FrameView(*TD.frame_view_args(frames[0], sf))
    rng = np.random.default_rng(2)
    for _ in range(200):
        x, y, r = float(rng.uniform(-30, S.KITTI_W + 30)), float(rng.uniform(-30, S.KITTI_H + 30)), float(rng.uniform(1, 80))
        lo, hi = int(rng.integers(-1, 5)), int(rng.integers(-1, 8))
        got, ref = both(oracle.features_in_area, fv, x, y, r, lo, hi)
        assert got.tolist() == ref.tolist()                      # same candidates in the same order


@pytest.mark.parametrize("th,mono,check", [(15.0, False, True), (7.0, False, True), (15.0, True, True), (30.0, False, False)])
def test_search_by_projection_last(seq_frames, th, mono, check):
    seq, frames, sf = seq_frames
    rng = np.random.default_rng(3)
    for (a, b) in ((0, 1), (1, 2)):
        last, cur = frames[a], frames[b]
        for trial in range(3):
            last_pose = seq.pose(a) if trial == 0 else (seq.pose(a) + np.r_[rot_pose(rng, 0.05, 0.003)[:4] * [1, 1, 1, 0], 0, 0, 0]).astype(np.float32)
            last_pose[:4] /= np.linalg.norm(last_pose[:4])
            cur_pose = seq.pose(b) if trial == 0 else (last_pose + np.r_[0, 0, 0, 0, rng.normal(0, 0.2, 3)]).astype(np.float32)
            xw, ok = TD.chain_unproject(last, last_pose)
            valid = (ok & (rng.random(len(ok)) < 0.9)).astype(np.uint8)
            obs_pos = (rng.random(len(ok)) < 0.8).astype(np.uint8)
            cur_state = rng.choice([0, 0, 0, 1, 2], len(cur["k"])).astype(np.uint8)
            fv = oracle.FrameView(*TD.frame_view_args(cur, sf))
            (n0, m0), (n1, m1) = both(oracle.search_by_projection_last, fv, cur_pose, last_pose, valid, xw, last["d"], last["k"]["octave"],
                                      last["k"]["angle"], obs_pos, th, mono, check, cur_state)
            # a free slot that was assigned and then cleared by the rotation check is NULL again: the reference cannot tell it from untouched
            m0n = np.where((m0 == -2) & (cur_state == 0), -1, m0)
            assert n0 == n1 and (m0n == m1).all(), (a, b, trial, n0, n1, int((m0n != m1).sum()))
            assert trial > 0 or n0 > 100


def test_unproject_stereo(seq_frames):
    """Frame::UnprojectStereo (the map points of the next frame's search) = the float32 glue the tests and the resident chain use."""
    seq, frames, sf = seq_frames
    rng = np.random.default_rng(4)
    fx, fy, cx, cy, bf = TD.CAM
    for fr in frames:
        for trial in range(3):
            pose = rot_pose(rng, 2.0, 0.05 * trial)
            pose[:4] /= np.linalg.norm(pose[:4])
            xw, ok = TD.chain_unproject(fr, pose)
            rx, rok = oracle.ref_unproject_stereo(pose, np.stack([fr["k"]["x"], fr["k"]["y"]], 1), fr["depth"], fx, fy, cx, cy)
            assert (ok == rok).all()
            assert (xw[ok] == rx[ok]).all(), np.abs(xw[ok] - rx[ok]).max()


def test_is_in_frustum_and_search_local(seq_frames):
    seq, frames, sf = seq_frames
    rng = np.random.default_rng(5)
    xw, desc, normal, mn, mx = TD.local_map(frames[:2], [seq.pose(0), seq.pose(1)], sf, rng)
    cur = frames[2]
    fv = oracle.FrameView(*TD.frame_view_args(cur, sf))
    for trial in range(4):
        pose = seq.pose(2).copy()
        if trial:
            pose = (pose + np.r_[rot_pose(rng, 0.0, 0.01)[:4] * [1, 1, 1, 0], rng.normal(0, 0.3, 3)]).astype(np.float32)
            pose[:4] /= np.linalg.norm(pose[:4])
        # Rcw, tcw, Ow as Frame::UpdatePoseMatrices derives them: taken from the reference's own SE3f through a 1-point unprojection
        q = pose[:4].astype(np.float64)
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]).astype(np.float32)
        tcw = pose[4:7]
        Ow = (-(R.T.astype(np.float64) @ tcw.astype(np.float64))).astype(np.float32)
        for cos_limit in (0.5, 0.9):
            t0, t1 = both(oracle.is_in_frustum, fv, R, tcw, Ow, xw, normal, mn, mx, cos_limit)
            for key in t0:
                assert (t0[key] == t1[key]).all(), (trial, key, int((t0[key] != t1[key]).sum()))
        tr = t0
        assert tr["in_view"].sum() > 200
        obs_pos = (rng.random(len(xw)) < 0.9).astype(np.uint8)
        for th, far, th_far in ((1.0, False, 0.0), (3.0, False, 0.0), (5.0, True, 25.0)):
            cs = rng.choice([0, 0, 1, 2], len(cur["k"])).astype(np.uint8)
            (n0, m0), (n1, m1) = both(oracle.search_by_projection_local, fv, tr, desc, obs_pos, th, 0.8, far, th_far, cs)
            assert n0 == n1 and (m0 == m1).all(), (trial, th, n0, n1)


def test_search_by_bow(seq_frames):
    seq, frames, sf = seq_frames
    rng = np.random.default_rng(6)
    kf, f = frames[0], frames[1]
    for bits in (4, 6):
        kcsr = TD.pseudo_feature_vector(kf["d"], bits); fcsr = TD.pseudo_feature_vector(f["d"], bits)
        valid = (rng.random(len(kf["d"])) < 0.85).astype(np.uint8)
        for ratio, check in ((0.7, True), (0.9, False)):
            (n0, m0), (n1, m1) = both(oracle.search_by_bow, kf["d"], kf["k"]["angle"], valid, kcsr, f["d"], f["k"]["angle"], fcsr, ratio, check)
            assert n0 == n1 and (m0 == m1).all()
            assert n0 > 50


def test_search_by_projection_reloc(seq_frames):
    seq, frames, sf = seq_frames
    rng = np.random.default_rng(7)
    kf, cur = frames[0], frames[1]
    xw, ok = TD.chain_unproject(kf, seq.pose(0))
    dist = np.linalg.norm(xw + seq.pose(0)[4:7], axis=1).astype(np.float32)
    mx = (dist * sf[kf["k"]["octave"]]).astype(np.float32); mn = (mx / sf[7]).astype(np.float32)
    fv = oracle.FrameView(*TD.frame_view_args(cur, sf))
    for th, orb in ((10.0, 100), (3.0, 64)):
        valid = (ok & (rng.random(len(ok)) < 0.9)).astype(np.uint8)
        occ = (rng.random(len(cur["k"])) < 0.2).astype(np.uint8)
        (n0, m0), (n1, m1) = both(oracle.search_by_projection_reloc, fv, seq.pose(1), valid, xw, kf["d"], kf["k"]["angle"], mn, mx, th, orb, True, occ)
        m0n = np.where((m0 == -2) & (occ == 0), -1, m0)
        assert n0 == n1 and (m0n == m1).all()


@pytest.mark.parametrize("seed", range(8))
def test_pose_optimization(seed):
    """Optimizer::PoseOptimization over the reference's g2o (LM, block solver, dense LDLT, Huber kernel, SE3Quat::exp) vs the restatement."""
    kw = [dict(), dict(n=300, outlier_frac=0.1), dict(n=1500, outlier_frac=0.4, stereo_frac=1.0), dict(n=500, stereo_frac=0.0),
          dict(n=40, outlier_frac=0.2), dict(n=9), dict(n=700, outlier_frac=0.5), dict(n=2000, outlier_frac=0.05, stereo_frac=0.5)][seed]
    p = TD.pose_problem(seed, **kw)
    rng = np.random.default_rng(seed)
    pose0 = p["pose0"] if seed % 2 == 0 else rot_pose(rng, 0.2, 0.01)
    (n0, pose_a, out_a), (n1, pose_b, out_b) = both(oracle.pose_optimize, pose0, p["xw"], p["obs"], p["inv_s2"], p["stereo"], *TD.CAM)
    assert n0 == n1 and (out_a == out_b).all(), (n0, n1, int((out_a != out_b).sum()))
    assert (pose_a == pose_b).all(), np.abs(pose_a - pose_b).max()         # observed: the float32 poses are bit-identical


def test_pose_optimization_degenerate():
    p = TD.pose_problem(0)
    for n in (0, 2, 3):
        (n0, a, oa), (n1, b, ob) = both(oracle.pose_optimize, p["pose0"], p["xw"][:n], p["obs"][:n], p["inv_s2"][:n], p["stereo"][:n], *TD.CAM)
        assert n0 == n1 and (oa == ob).all() and np.abs(a - b).max() <= 1e-6


def test_stereo_matches_and_rgbd():
    """Frame::ComputeStereoMatches on a rectified pair with a smooth non-uniform disparity field; Frame::ComputeStereoFromRGBD."""
    left, right = S.stereo_pair(11, 640, 376)
    exl, exr = oracle.Extractor(1000), oracle.Extractor(1000)
    kl, dl, _ = exl(left); kr, dr, _ = exr(right)
    mbf = S.KITTI_BF; mb = mbf / S.KITTI_FX
    (d0, u0), (d1, u1) = both(oracle.stereo_matches, kl, dl, kr, dr, exl, exr, mb, mbf)
    assert (d0 == d1).all() and (u0 == u1).all()
    assert (d0 > 0).sum() > 100
    rng = np.random.default_rng(1)
    dm = np.where(rng.random((376, 640)) < 0.6, rng.uniform(3, 80, (376, 640)), 0).astype(np.float32)
    xy = np.stack([kl["x"], kl["y"]], 1)
    rd, ru = oracle.ref_stereo_from_rgbd(xy, xy, dm, mbf)
    d = dm[kl["y"].astype(np.int32), kl["x"].astype(np.int32)]
    exp_d = np.where(d > 0, d, np.float32(-1)); exp_u = np.where(d > 0, kl["x"] - np.float32(mbf) / np.where(d > 0, d, 1).astype(np.float32), np.float32(-1)).astype(np.float32)
    assert (rd == exp_d).all() and (ru == exp_u).all()


def test_distinctive_descriptors():
    rng = np.random.default_rng(9)
    counts = rng.integers(0, 12, 60); start = np.r_[0, np.cumsum(counts)].astype(np.int32)
    base = rng.integers(0, 256, (60, 32), dtype=np.uint8)
    desc = np.concatenate([base[i] ^ (rng.random((c, 32)) < 0.06).astype(np.uint8) * rng.integers(0, 256, (c, 32), dtype=np.uint8) for i, c in enumerate(counts)] + [np.zeros((0, 32), np.uint8)])
    b0, b1 = both(oracle.distinctive_descriptors, start, desc)
    for i, c in enumerate(counts):
        if c == 0:
            continue
        assert (desc[start[i] + b0[i]] == desc[start[i] + b1[i]]).all()


def test_search_for_triangulation():
    """The reference's own SearchForTriangulation + Pinhole::epipolarConstrain vs the restatement, fed the F12 / epipole the reference
    derives from the two key-frame poses."""
    rng = np.random.default_rng(21)
    n1, n2, n_nodes = 600, 650, 30
    base = rng.integers(0, 256, (80, 32), dtype=np.uint8)

    def frame(n):
        which = rng.integers(0, len(base), n)
        d = base[which] ^ (rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8))
        k = np.zeros(n, oracle.KP_DTYPE)
        k["x"] = rng.uniform(0, 1241, n); k["y"] = rng.uniform(0, 376, n); k["octave"] = rng.integers(0, 8, n); k["angle"] = rng.uniform(0, 360, n)
        node = which % n_nodes
        order = np.argsort(node, kind="stable")
        ids, start = np.unique(node[order], return_index=True)
        fv = (ids.astype(np.uint32), np.r_[start, n].astype(np.int32), order.astype(np.int32))
        return dict(desc=d, keys=k, has_mp=rng.random(n) < 0.3, uright=np.where(rng.random(n) < 0.5, rng.uniform(0, 1000, n), -1).astype(np.float32), fv=fv)

    kf1, kf2 = frame(n1), frame(n2)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32); sg = (sf * sf).astype(np.float32)
    for trial in range(3):
        T1w = rot_pose(rng, 1.0, 0.05); T2w = rot_pose(rng, 1.0, 0.05)
        T1w[:4] /= np.linalg.norm(T1w[:4]); T2w[:4] /= np.linalg.norm(T2w[:4])
        for only_stereo, coarse, ori in ((0, 1, 1), (1, 1, 0), (0, 0, 1), (1, 0, 1)):
            n_ref, m_ref, F12, ep = oracle.ref_search_for_triangulation(T1w, T2w, TD.CAM, kf1, kf2, sf, sg, only_stereo, coarse, ori)
            n_or, m_or = oracle.search_for_triangulation(kf1, kf2, F12, ep, sf, sg, only_stereo, coarse, ori)
            assert n_ref == n_or and (m_ref == m_or).all(), (trial, only_stereo, coarse, ori, n_ref, n_or)
            if coarse:
                assert n_ref > 20


def test_fuse(seq_frames):
    """The reference's own ORBmatcher::Fuse (search + bookkeeping) vs the restated search: a point is fused into slot best_idx exactly
    when the restatement's best Hamming distance is <= TH_LOW (src/ORBmatcher.cc:1306-1325)."""
    seq, frames, sf = seq_frames
    rng = np.random.default_rng(22)
    xw, desc, normal, mn, mx = TD.local_map([frames[0]], [seq.pose(0)], sf, rng)
    kf = frames[1]
    fv = oracle.FrameView(*TD.frame_view_args(kf, sf))
    for trial, th in enumerate((3.0, 4.0)):
        pose = seq.pose(1).copy()
        if trial:
            pose = (pose + np.r_[rot_pose(rng, 0.0, 0.005)[:4] * [1, 1, 1, 0], rng.normal(0, 0.1, 3)]).astype(np.float32)
            pose[:4] /= np.linalg.norm(pose[:4])
        valid = (rng.random(len(xw)) < 0.9).astype(np.uint8)
        nf, bi_ref, Ow = oracle.ref_fuse(fv, pose, valid, xw, normal, mn, mx, desc, th)
        bi, bd = oracle.fuse_search(fv, pose, Ow, valid, xw, normal, mn, mx, desc, th)
        exp = np.where(bd <= 50, bi, -1)
        assert (exp == bi_ref).all(), (trial, int((exp != bi_ref).sum()))
        assert nf == int((exp >= 0).sum()) and nf > 100


@pytest.mark.parametrize("seed,kw", [(1, dict()), (2, dict(n_kf=12, n_fixed=3, n_points=900, outlier_frac=0.05)), (3, dict(n_kf=5, n_fixed=1, n_points=200, stereo_frac=1.0)),
                                     (4, dict(n_kf=6, n_fixed=2, n_points=300, stereo_frac=0.0))])
def test_local_bundle_adjustment(seed, kw):
    """Optimizer::LocalBundleAdjustment's own body (src/Optimizer.cc:1116-1499) on the reference's g2o (Schur-complement block solver,
    LM, EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ, Huber) vs the dense restatement.  The reference solves the reduced system with a
    sparse Cholesky (stood in for by the dense LDLT here) and walks the edges in map order, so the comparison is tolerance-based:
    poses and points to 1e-4, identical erase decisions."""
    import ba_data as D
    p = D.make_problem(seed, **kw)
    (po0, pt0, er0, it0, _), (po1, pt1, er1, _, _) = both(oracle.local_bundle_adjustment, *D.args(p))
    assert np.abs(po0 - po1).max() < 1e-4, np.abs(po0 - po1).max()
    assert np.abs(pt0 - pt1).max() < 1e-4, np.abs(pt0 - pt1).max()
    assert (er0 == er1).all(), int((er0 != er1).sum())
    assert np.array_equal(po1[p["pose_fixed"] != 0], p["poses"][p["pose_fixed"] != 0]) or np.abs(po1[p["pose_fixed"] != 0] - p["poses"][p["pose_fixed"] != 0]).max() < 1e-6
