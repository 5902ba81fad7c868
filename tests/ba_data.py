"""Synthetic local-BA problems (flat graph as the C-ABI shim gathers it from Optimizer::LocalBundleAdjustment, src/Optimizer.cc:1116-1404):
key-frame poses along a forward trajectory with yaw, points in front of them, mono / stereo observations with pixel noise per octave,
a few gross outliers, perturbed initial estimates; the first key frames are fixed."""
import numpy as np

from orb_slam3_rgbl_b200 import synthetic as S


def _quat_yaw(a):
    return np.array([0.0, np.sin(a / 2), 0.0, np.cos(a / 2)])       # rotation about the camera y axis (x, y, z, w)


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_problem(seed, n_kf=8, n_fixed=2, n_points=600, outlier_frac=0.03, stereo_frac=0.6, pose_noise=(0.004, 0.03), point_noise=0.06):
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, bf = S.KITTI_FX, S.KITTI_FY, S.KITTI_CX, S.KITTI_CY, S.KITTI_BF
    W, H = S.KITTI_W, S.KITTI_H
    # true poses Tcw: camera moves forward (world z) with a slow yaw
    true_poses = []
    for k in range(n_kf):
        q = _quat_yaw(0.01 * k)
        cpos = np.array([0.05 * k, 0.0, 0.9 * k])                    # camera centre in the world
        R = _rot(q)
        t = -R @ cpos
        true_poses.append(np.concatenate([q, t]))
    true_poses = np.array(true_poses)
    pts = np.column_stack([rng.uniform(-12, 12, n_points), rng.uniform(-2.5, 1.5, n_points), rng.uniform(6, 45, n_points) + 0.9 * n_kf * rng.random(n_points)])
    e_point, e_pose, obs, stereo, inv_s2, is_out = [], [], [], [], [], []
    for j in range(n_points):
        for k in range(n_kf):
            R = _rot(true_poses[k, :4]); pc = R @ pts[j] + true_poses[k, 4:]
            if pc[2] < 1.0:
                continue
            u = fx * pc[0] / pc[2] + cx; v = fy * pc[1] / pc[2] + cy
            if not (20 < u < W - 20 and 20 < v < H - 20) or rng.random() < 0.15:
                continue
            octave = int(rng.integers(0, 8)); sig = 1.2 ** octave
            st = rng.random() < stereo_frac
            out = rng.random() < outlier_frac
            nu, nv = rng.normal(0, 0.6 * sig, 2)
            if out:
                nu += rng.choice([-1, 1]) * rng.uniform(15, 60); nv += rng.choice([-1, 1]) * rng.uniform(10, 40)
            ur = (u - bf / pc[2] + rng.normal(0, 0.6 * sig)) if st else -1.0
            if ur < 0:                 # the reference tells a stereo observation by mvuRight >= 0 (src/Optimizer.cc:1286-1311)
                st, ur = False, -1.0
            e_point.append(j); e_pose.append(k); obs.append([u + nu, v + nv, ur]); stereo.append(st); inv_s2.append(1.0 / (sig * sig)); is_out.append(out)
    # Optimizer::LocalBundleAdjustment only adjusts points seen from a LOCAL (non-fixed) key frame (src/Optimizer.cc:1133-1160): drop the
    # others, as the gathering shim would
    e_point = np.array(e_point, np.int64); e_pose = np.array(e_pose, np.int64)
    local = np.zeros(n_points, bool); local[e_point[e_pose >= n_fixed]] = True
    keep_e = local[e_point]
    remap = np.cumsum(local) - 1
    e_point = remap[e_point[keep_e]]; e_pose = e_pose[keep_e]
    obs = [o for o, k in zip(obs, keep_e) if k]; stereo = [o for o, k in zip(stereo, keep_e) if k]
    inv_s2 = [o for o, k in zip(inv_s2, keep_e) if k]; is_out = [o for o, k in zip(is_out, keep_e) if k]
    pts = pts[local]
    init_poses = true_poses.copy()
    for k in range(n_fixed, n_kf):
        dq = np.concatenate([rng.normal(0, pose_noise[0], 3), [1.0]]); dq /= np.linalg.norm(dq)
        x1, y1, z1, w1 = dq; x2, y2, z2, w2 = init_poses[k, :4]
        q = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
        init_poses[k, :4] = q / np.linalg.norm(q)
        init_poses[k, 4:] += rng.normal(0, pose_noise[1], 3)
    init_pts = pts + rng.normal(0, point_noise, pts.shape)
    fixed = np.zeros(n_kf, np.uint8); fixed[:n_fixed] = 1
    return dict(poses=init_poses.astype(np.float32), pose_fixed=fixed, points=init_pts.astype(np.float32),
                e_point=np.array(e_point, np.int32), e_pose=np.array(e_pose, np.int32), obs=np.array(obs, np.float32),
                stereo=np.array(stereo, np.uint8), inv_sigma2=np.array(inv_s2, np.float32), cam=(fx, fy, cx, cy, bf),
                true_poses=true_poses, true_points=pts, is_outlier=np.array(is_out, bool))


def args(p):
    return (p["poses"], p["pose_fixed"], p["points"], p["e_point"], p["e_pose"], p["obs"], p["stereo"], p["inv_sigma2"], *p["cam"])
