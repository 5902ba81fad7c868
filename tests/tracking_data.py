"""Shared builders of matcher / pose-optimisation inputs from the synthetic plane sequence (CPU oracle only)."""
import numpy as np

import oracle
from orb_slam3_rgbl_b200 import synthetic as S

CAM = (S.KITTI_FX, S.KITTI_FY, S.KITTI_CX, S.KITTI_CY, S.KITTI_BF)


def extract_frames(seq, ts, nfeatures=2000):
    ex = oracle.Extractor(nfeatures)
    mask = S.structuring_element("diamond", 5)
    out = []
    for t in ts:
        img, pts = seq.image(t), seq.cloud(t)
        k, d, _ = ex(img)
        dep, ur, _, _ = oracle.depth_from_pcd(pts, seq.P, seq.W, seq.H, mask, S.KITTI_BF, k, k)
        out.append(dict(t=t, k=k, d=d, depth=dep, ur=ur))
    return out, ex.scale_factors.copy()


def unproject(fr, pose):
    """Frame::UnprojectStereo for every keypoint with depth (src/Frame.cc:1137-1150): world points for an identity-rotation pose."""
    k, z = fr["k"], fr["depth"]
    f32 = np.float32
    zz = np.where(z > 0, z, f32(1)).astype(f32)
    x = ((k["x"] - f32(S.KITTI_CX)) * zz * f32(1.0 / np.float32(S.KITTI_FX))).astype(f32)
    y = ((k["y"] - f32(S.KITTI_CY)) * zz * f32(1.0 / np.float32(S.KITTI_FY))).astype(f32)
    xc = np.stack([x, y, zz], 1)
    twc = -np.asarray(pose[4:7], f32)            # R = I
    return (xc + twc).astype(f32), (z > 0)


def frame_view_args(fr, sf):
    return (fr["k"], fr["ur"], fr["d"], S.KITTI_W, S.KITTI_H, sf) + CAM


def local_map(frames, poses, sf, rng, n_levels=8):
    """A small 'local map': world points of several frames with normals / scale-invariance distances
    (MapPoint::UpdateNormalAndDepth, src/MapPoint.cc:437-490)."""
    xs, ds, ns, mins, maxs = [], [], [], [], []
    for fr, pose in zip(frames, poses):
        xw, ok = unproject(fr, pose)
        ow = -np.asarray(pose[4:7], np.float32)
        po = xw[ok] - ow
        dist = np.linalg.norm(po, axis=1).astype(np.float32)
        mx = (dist * sf[fr["k"]["octave"][ok]]).astype(np.float32)
        xs.append(xw[ok]); ds.append(fr["d"][ok]); ns.append((po / dist[:, None]).astype(np.float32))
        maxs.append(mx); mins.append((mx / sf[n_levels - 1]).astype(np.float32))
    xw = np.concatenate(xs); perm = rng.permutation(len(xw))
    return (xw[perm], np.concatenate(ds)[perm], np.concatenate(ns)[perm], np.concatenate(mins)[perm], np.concatenate(maxs)[perm])


def pose_problem(seed, n=900, outlier_frac=0.3, stereo_frac=0.7):
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, bf = CAM

    def quat(w):
        th = np.linalg.norm(w)
        return np.r_[w / th * np.sin(th / 2), np.cos(th / 2)] if th > 0 else np.array([0, 0, 0, 1.0])

    def qrot(q, p):
        v = q[:3]; uv = 2 * np.cross(v, p)
        return p + q[3] * uv + np.cross(v, uv)

    q_true = quat(rng.normal(0, 0.02, 3)); t_true = rng.normal(0, 0.4, 3)
    Pc = np.stack([rng.uniform(-20, 20, n), rng.uniform(-3, 3, n), rng.uniform(6, 60, n)], 1)
    qc = q_true * np.array([-1, -1, -1, 1])
    Xw = np.array([qrot(qc, p - t_true) for p in Pc])
    u = fx * Pc[:, 0] / Pc[:, 2] + cx; v = fy * Pc[:, 1] / Pc[:, 2] + cy; ur = u - bf / Pc[:, 2]
    lvl = rng.integers(0, 8, n); sig = 1.2 ** lvl
    obs = np.stack([u + rng.normal(0, 0.5, n) * sig, v + rng.normal(0, 0.5, n) * sig, ur + rng.normal(0, 0.5, n) * sig], 1)
    stereo = ((rng.random(n) < stereo_frac) & (obs[:, 2] >= 0)).astype(np.uint8); obs[stereo == 0, 2] = -1     # the reference tells stereo by mvuRight >= 0
    oi = rng.choice(n, int(outlier_frac * n), replace=False)
    obs[oi, :2] += rng.uniform(-60, 60, (len(oi), 2))
    inv_s2 = (1.0 / sig ** 2).astype(np.float32)
    pose0 = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    return dict(pose0=pose0, xw=Xw.astype(np.float32), obs=obs.astype(np.float32), inv_s2=inv_s2, stereo=stereo,
                truth=np.r_[q_true, t_true])


# ---- the chain glue lives in oracle/chain.py (shared with bench.py's CPU legs); KITTI-size wrappers for the tests ----
from oracle import chain as _chain  # noqa: E402
from oracle.chain import se3f_rotate, se3f_inverse, quat_to_matrix_f32, pose_matrices  # noqa: E402,F401

f32 = np.float32


def chain_unproject(fr, pose):
    return _chain.chain_unproject(fr, pose, CAM)


def local_points_of(fr, pose, sf):
    return _chain.local_points_of(fr, pose, sf, CAM)


def oracle_chain(frames, sf, pose0, th=15.0):
    return _chain.oracle_chain(frames, sf, pose0, S.KITTI_W, S.KITTI_H, CAM, th)


def oracle_chain2(frames, sf, pose0, K=3, th_last=15.0, th_local=3.0, nn_ratio=0.8, state=None):
    return _chain.oracle_chain2(frames, sf, pose0, S.KITTI_W, S.KITTI_H, CAM, K, th_last, th_local, nn_ratio, state)


def pseudo_feature_vector(desc, n_bits=6):
    """Stand-in for DBoW2's FeatureVector (node -> feature indices): node id = a few stable descriptor bits.  Returns the
    CSR triple (ascending node ids, node_start, feature indices in ascending feature order, like Frame::ComputeBoW)."""
    node = (desc[:, 0].astype(np.uint32) >> (8 - n_bits)) * 7 + 3           # arbitrary non-contiguous ids
    ids = np.unique(node)
    start = [0]; feat = []
    for n in ids:
        idx = np.nonzero(node == n)[0]
        feat.extend(idx.tolist()); start.append(len(feat))
    return ids.astype(np.uint32), np.array(start, np.int32), np.array(feat, np.int32)
