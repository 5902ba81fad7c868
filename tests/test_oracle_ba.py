"""LocalBundleAdjustment oracle (oracle/ba_oracle.cpp): convergence and invariants on synthetic graphs.  CPU only.  The
restatement cannot be pinned against g2o here (no Eigen), so it is held to what the algorithm must do."""
import numpy as np

import oracle
import ba_data as D


def test_noise_free_graph_converges_to_the_truth():
    p = D.make_problem(2, outlier_frac=0.0)
    fx, fy, cx, cy, bf = p["cam"]
    obs = p["obs"].copy()
    for k, (j, i) in enumerate(zip(p["e_point"], p["e_pose"])):
        pc = D._rot(p["true_poses"][i, :4]) @ p["true_points"][j] + p["true_poses"][i, 4:]
        u = fx * pc[0] / pc[2] + cx; v = fy * pc[1] / pc[2] + cy
        obs[k] = [u, v, (u - bf / pc[2]) if p["stereo"][k] else -1]
    p["obs"] = obs.astype(np.float32)
    po, pt, er, it, chi = oracle.local_bundle_adjustment(*D.args(p))
    assert chi < 0.05 and er.sum() == 0
    assert np.abs(po[:, 4:] - p["true_poses"][:, 4:]).max() < 2e-3
    assert np.array_equal(po[:2], p["poses"][:2])                       # fixed key frames untouched


def test_outliers_are_flagged_and_chi2_decreases():
    p = D.make_problem(1)
    chis = [oracle.local_bundle_adjustment(*D.args(p), iterations=k)[4] for k in (1, 2, 5, 10)]
    assert chis[0] > chis[1] > chis[2] > chis[3]
    po, pt, er, it, chi = oracle.local_bundle_adjustment(*D.args(p))
    assert er[p["is_outlier"]].all() and er[~p["is_outlier"]].mean() < 0.01
    assert np.abs(po[:, 4:] - p["true_poses"][:, 4:]).max() < 0.08


def test_degenerate_graphs():
    p = D.make_problem(3, n_kf=4, n_points=50)
    a = list(D.args(p))
    a[1] = np.ones_like(p["pose_fixed"])                                # every key frame fixed: structure-only adjustment
    po, pt, er, it, chi = oracle.local_bundle_adjustment(*a)
    assert np.array_equal(po, p["poses"]) and it >= 1 and np.isfinite(pt).all()
    none = (p["poses"], p["pose_fixed"], p["points"], np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3), np.float32),
            np.zeros(0, np.uint8), np.zeros(0, np.float32), *p["cam"])
    po, pt, er, it, chi = oracle.local_bundle_adjustment(*none)
    assert it == 0 and np.array_equal(po, p["poses"]) and np.array_equal(pt, p["points"])
