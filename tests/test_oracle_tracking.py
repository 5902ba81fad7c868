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
