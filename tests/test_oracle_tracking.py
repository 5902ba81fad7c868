"""CPU checks of the matcher / pose oracle (properties + consistency; there are no reference goldens for these)."""
import numpy as np
import pytest

import oracle
import tracking_data as TD
from orb_slam3_rgbl_b200 import synthetic as S


@pytest.fixture(scope="module")
def seq_frames():
    seq = S.PlaneSequence(5, 3)
    frames, sf = TD.extract_frames(seq, [0, 1])
    return seq, frames, sf


def test_descriptor_distance_is_popcount():
    rng = np.random.default_rng(1)
    for _ in range(100):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert oracle.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())


def test_features_in_area_matches_bruteforce(seq_frames):
    seq, frames, sf = seq_frames
    fv = oracle.FrameView(*TD.frame_view_args(frames[0], sf))
    k = frames[0]["k"]
    rng = np.random.default_rng(2)
    for _ in range(50):
        x, y, r = float(rng.uniform(0, S.KITTI_W)), float(rng.uniform(0, S.KITTI_H)), float(rng.uniform(2, 60))
        lo, hi = int(rng.integers(-1, 4)), int(rng.integers(-1, 8))
        got = oracle.features_in_area(fv, x, y, r, lo, hi)
        ok = (np.abs(k["x"] - np.float32(x)) < np.float32(r)) & (np.abs(k["y"] - np.float32(y)) < np.float32(r))
        if lo > 0 or hi >= 0:
            ok &= k["octave"] >= lo
            if hi >= 0:
                ok &= k["octave"] <= hi
        assert sorted(got.tolist()) == np.nonzero(ok)[0].tolist()


def test_search_last_recovers_the_shift(seq_frames):
    seq, frames, sf = seq_frames
    last, cur = frames
    xw, ok = TD.unproject(last, seq.pose(0))
    n, match = oracle.search_by_projection_last(oracle.FrameView(*TD.frame_view_args(cur, sf)), seq.pose(1), seq.pose(0), ok.astype(np.uint8), xw,
                                                last["d"], last["k"]["octave"], last["k"]["angle"], np.ones(len(ok), np.uint8), 15.0)
    m = np.nonzero(match >= 0)[0]
    assert n == len(m) and n > 200
    assert len(set(match[m].tolist())) == len(m)                 # a map point is assigned at most once
    dx = cur["k"]["x"][m] - last["k"]["x"][match[m]]
    assert np.mean(np.abs(dx + seq.shift * sf[cur["k"]["octave"][m]] / sf[last["k"]["octave"][match[m]]]) < 3) > 0.8


def test_pose_oracle_converges_and_flags_outliers():
    p = TD.pose_problem(0)
    n, pose, out = oracle.pose_optimize(p["pose0"], p["xw"], p["obs"], p["inv_s2"], p["stereo"], *TD.CAM)
    assert np.abs(pose[4:] - p["truth"][4:]).max() < 0.05 and np.abs(pose[:4] - p["truth"][:4]).max() < 2e-3
    assert n == len(out) - out.sum() and 0.25 < out.mean() < 0.4
    n2, pose2, _ = oracle.pose_optimize(p["pose0"], p["xw"][:2], p["obs"][:2], p["inv_s2"][:2], p["stereo"][:2], *TD.CAM)
    assert n2 == 0 and (pose2 == p["pose0"]).all()                  # < 3 correspondences (src/Optimizer.cc:996)


def test_pose_oracle_result_is_a_stationary_point_of_the_stated_cost():
    """Independent check of the LM restatement: the 4th optimisation round runs WITHOUT the robust kernel on the inlier set it
    ends with (src/Optimizer.cc:1004-1012), so the returned pose must zero the gradient of sum_e e^T Omega e over the edges not
    flagged as outliers.  The cost is restated here in numpy (float64) with a numeric SE3 perturbation - nothing shared with
    oracle/pose_oracle.cpp."""
    fx, fy, cx, cy, bf = TD.CAM

    def qmul(a, b):
        ax, ay, az, aw = a; bx, by, bz, bw = b
        return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
                         aw * bw - ax * bx - ay * by - az * bz])

    def qrot(q, P):
        v = q[:3]; uv = 2 * np.cross(v, P)
        return P + q[3] * uv + np.cross(v, uv)

    def cost(pose, pr, inl):
        Pc = qrot(pose[:4], pr["xw"][inl].astype(np.float64)) + pose[4:]
        u = fx * Pc[:, 0] / Pc[:, 2] + cx; v = fy * Pc[:, 1] / Pc[:, 2] + cy; ur = u - bf / Pc[:, 2]
        o = pr["obs"][inl].astype(np.float64); st = pr["stereo"][inl] != 0; w = pr["inv_s2"][inl].astype(np.float64)
        e2 = (o[:, 0] - u) ** 2 + (o[:, 1] - v) ** 2 + np.where(st, (o[:, 2] - ur) ** 2, 0.0)
        return float((w * e2).sum())

    for seed in (0, 3, 7):
        pr = TD.pose_problem(seed, n=700, outlier_frac=0.2)
        n_in, pose, out = oracle.pose_optimize(pr["pose0"], pr["xw"], pr["obs"], pr["inv_s2"], pr["stereo"], *TD.CAM)
        inl = out == 0
        pose = pose.astype(np.float64); pose[:4] /= np.linalg.norm(pose[:4])
        c0 = cost(pose, pr, inl)
        g = np.zeros(6); h = 1e-6
        for i in range(6):
            d = np.zeros(6); d[i] = h
            def moved(sgn):
                w = sgn * d[:3]; th = np.linalg.norm(w)
                dq = np.r_[w / th * np.sin(th / 2), np.cos(th / 2)] if th > 0 else np.array([0, 0, 0, 1.0])
                return np.r_[qmul(dq, pose[:4]), qrot(dq, pose[4:]) + sgn * d[3:]]
            g[i] = (cost(moved(+1), pr, inl) - cost(moved(-1), pr, inl)) / (2 * h)
        # scale: a 1 cm / 0.01 rad step must change the cost by far more than the gradient predicts (float32 pose output limits the zero)
        step = np.array([1e-2] * 3 + [1e-2] * 3)
        assert np.abs(g * step).max() < 2e-3 * max(c0, 1.0), (seed, g, c0)


def test_fuse_search_oracle_vs_bruteforce(seq_frames):
    """orc_fuse_search against a numpy restatement of ORBmatcher::Fuse's search loop (src/ORBmatcher.cc:1176-1303) with a
    brute-force window scan instead of the grid."""
    seq, frames, sf = seq_frames
    rng = np.random.default_rng(4)
    xw, desc, normal, mn, mx = TD.local_map([frames[0]], [seq.pose(0)], sf, rng)
    kf = frames[1]; pose = seq.pose(1); Ow = -pose[4:7]
    fv = oracle.FrameView(*TD.frame_view_args(kf, sf))
    valid = (rng.random(len(xw)) < 0.9).astype(np.uint8)
    th = 3.0
    bi, bd = oracle.fuse_search(fv, pose, Ow, valid, xw, normal, mn, mx, desc, th)
    fx, fy, cx, cy, bf = TD.CAM
    f32 = np.float32
    pop = np.array([bin(i).count("1") for i in range(256)], np.int32)
    k = kf["k"]; ur = kf["ur"]
    checked = 0
    for i in range(0, len(xw), 7):
        exp = (-1, 256)
        if valid[i]:
            Pc = (xw[i] + pose[4:7]).astype(f32)                  # identity rotation in the synthetic sequence
            if Pc[2] >= 0:
                u = f32(f32(f32(fx) * Pc[0]) / Pc[2]) + f32(cx); v = f32(f32(f32(fy) * Pc[1]) / Pc[2]) + f32(cy)
                if 0 <= u < S.KITTI_W and 0 <= v < S.KITTI_H:
                    PO = (xw[i] - Ow).astype(f32); d3 = f32(np.sqrt(f32(f32(PO[0] * PO[0] + PO[1] * PO[1]) + PO[2] * PO[2])))
                    if not (d3 < f32(0.8) * mn[i] or d3 > f32(1.2) * mx[i]) and not (float(f32(f32(PO[0] * normal[i][0] + PO[1] * normal[i][1]) + PO[2] * normal[i][2])) < 0.5 * float(d3)):
                        level = int(np.ceil(np.log(f32(mx[i] / d3)) / f32(np.log(1.2))))
                        level = min(max(level, 0), 7)
                        r = f32(th) * sf[level]; urp = u - f32(f32(bf) * f32(1 / Pc[2]))
                        best = (256, -1)
                        # candidates in the reference's order are irrelevant for a strict '<' minimum up to ties: compare sets of minima
                        cand = []
                        for j in range(len(k)):
                            if not (abs(k["x"][j] - u) < r and abs(k["y"][j] - v) < r):
                                continue
                            if k["octave"][j] < level - 1 or k["octave"][j] > level:
                                continue
                            inv_s2 = f32(1) / f32(sf[k["octave"][j]] * sf[k["octave"][j]])
                            ex = u - k["x"][j]; ey = v - k["y"][j]
                            if ur[j] >= 0:
                                er = urp - ur[j]
                                if float(f32(f32(f32(ex * ex + ey * ey) + er * er) * inv_s2)) > 7.8:
                                    continue
                            elif float(f32(f32(ex * ex + ey * ey) * inv_s2)) > 5.99:
                                continue
                            cand.append((int(pop[desc[i] ^ kf["d"][j]].sum()), j))
                        if cand:
                            dmin = min(c[0] for c in cand)
                            exp = ({j for d, j in cand if d == dmin}, dmin)
        if exp[0] == -1:
            assert bi[i] == -1 and bd[i] == 256
        else:
            assert bd[i] == exp[1] and bi[i] in exp[0]
            checked += 1
    assert checked > 20


def test_motion_model_pose_prediction():
    """oracle.chain.se3f_mul / predict_pose restate Sophus::SE3f's product (Thirdparty/Sophus/sophus/se3.hpp:304-308, so3.hpp:325-339 + the
    normalising quaternion constructor) and Tracking's constant-velocity model (src/Tracking.cc:2243-2245, 2904): group laws in float32
    tolerance, unit quaternions, and exact extrapolation of a constant motion."""
    from oracle import chain as OC
    rng = np.random.default_rng(5)

    def rand_pose():
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        return np.concatenate([q, rng.normal(size=3) * 3]).astype(np.float32)

    def mat(p):
        R = OC.quat_to_matrix_f32(p[:4]).astype(np.float64); T = np.eye(4); T[:3, :3] = R; T[:3, 3] = p[4:]
        return T

    for _ in range(50):
        a, b = rand_pose(), rand_pose()
        ab = OC.se3f_mul(a, b)
        assert abs(np.linalg.norm(ab[:4].astype(np.float64)) - 1) < 2e-7
        assert np.allclose(mat(ab), mat(a) @ mat(b), atol=2e-5)
        qi, ti = OC.se3f_inverse(a)
        assert np.allclose(mat(OC.se3f_mul(a, np.concatenate([qi, ti]).astype(np.float32))), np.eye(4), atol=2e-5)
        # constant motion M: T1 = M T0, so the prediction for the next frame is M T1
        M = rand_pose(); M[4:] *= 0.05; M[:3] *= 0.02; M[:4] /= np.linalg.norm(M[:4])
        T0 = rand_pose(); T1 = OC.se3f_mul(M, T0)
        assert np.allclose(mat(OC.predict_pose(T0, T1)), mat(M) @ mat(T1), atol=5e-5)
    assert (OC.predict_pose(None, a) == a).all()


def test_oracle_chain_is_the_same_in_one_piece_and_in_batches():
    """oracle.chain.oracle_chain2 is what the GPU chain tests and bench.py's CPU arms compare / time against.  Its `state` (last frame, its
    pose, the pose before it for the constant-velocity prediction, the local-map ring and its frame counter) must carry a sequence from one
    batch into the next exactly: 7 frames tracked in one call, in batches of 3 + 2 + 2, and one frame at a time give bitwise the same
    poses and counts; and the prediction is used (the first optimisation of a frame starts closer to its result than the last pose is)."""
    n = 7
    seq = S.PlaneSequence(83, n + 1)
    frames, sf = TD.extract_frames(seq, list(range(n)), nfeatures=1000)
    whole = TD.oracle_chain2(frames, sf, seq.pose(0), K=2)
    for cuts in ([3, 5], [1, 2, 3, 4, 5, 6]):
        state, parts = None, []
        for a, b in zip([0] + cuts, cuts + [n]):
            out = TD.oracle_chain2(frames[a:b], sf, seq.pose(0), K=2, state=state)
            state = out[-1]
            parts.append(out[:-1])
        for i in range(5):
            got = np.concatenate([p[i] for p in parts])
            assert got.shape == whole[i].shape and (got == whole[i]).all(), (cuts, i)
    poses = whole[0]
    from oracle import chain as OC
    for t in range(2, n):
        pred = OC.predict_pose(poses[t - 2], poses[t - 1])
        assert np.abs(pred - poses[t]).max() < np.abs(poses[t - 1] - poses[t]).max(), t
    assert whole[3][3:].min() > 20                     # the local map contributes matches once the ring holds frames
