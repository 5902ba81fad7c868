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
