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


# ---- float32 mirror of the resident chain's glue (Frame::UnprojectStereo with a Sophus SE3f pose) ----
f32 = np.float32


def se3f_rotate(q, p):
    """Sophus SO3f * point (so3.hpp:358-366), float32, same operation order as the kernels."""
    qx, qy, qz, qw = (f32(v) for v in q[:4])
    px, py, pz = p[:, 0], p[:, 1], p[:, 2]
    uv = np.stack([qy * pz - qz * py, qz * px - qx * pz, qx * py - qy * px], 1).astype(f32)
    uv = (uv + uv).astype(f32)
    c = np.stack([qy * uv[:, 2] - qz * uv[:, 1], qz * uv[:, 0] - qx * uv[:, 2], qx * uv[:, 1] - qy * uv[:, 0]], 1).astype(f32)
    return ((p + qw * uv).astype(f32) + c).astype(f32)


def se3f_inverse(pose):
    """Sophus::SE3f::inverse() (se3.hpp:208-211): invR = SO3f(conjugate) - the quaternion constructor normalises in float
    (so3.hpp:229-231, 481-487) - and translation invR * (t * -1).  -> (q_inv[4], t_inv[3]) float32."""
    q = np.array([-pose[0], -pose[1], -pose[2], pose[3]], f32)
    length = np.sqrt(f32(f32(q[0] * q[0]) + f32(q[1] * q[1])) + f32(f32(q[2] * q[2]) + f32(q[3] * q[3])))       # Eigen: (x2 + y2) + (z2 + w2)
    q = (q / f32(length)).astype(f32)
    nt = (np.asarray(pose[4:7], f32) * f32(-1.0)).astype(f32)[None, :]
    return q, se3f_rotate(q, nt)[0]


def quat_to_matrix_f32(q):
    """Eigen QuaternionBase::toRotationMatrix in float32 (Geometry/Quaternion.h)."""
    x, y, z, w = (f32(v) for v in q)
    tx, ty, tz = f32(2) * x, f32(2) * y, f32(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[f32(1) - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, f32(1) - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, f32(1) - (txx + tyy)]], f32)


def chain_unproject(fr, pose):
    """Frame::UnprojectStereo (src/Frame.cc:1137-1150) for every keypoint: x3D = mRwc * x3Dc + mOw with mRwc / mOw from
    Frame::UpdatePoseMatrices (src/Frame.cc:562-569), float32, Eigen 3.3's product order r0*x + (r1*y + r2*z).
    Pinned to the reference's own code by tests/test_oracle_tracking_ref.py::test_unproject_stereo."""
    k, z = fr["k"], fr["depth"]
    ok = z > 0
    zz = np.where(ok, z, f32(1)).astype(f32)
    invfx, invfy = f32(1.0) / f32(S.KITTI_FX), f32(1.0) / f32(S.KITTI_FY)
    pc = np.stack([((k["x"] - f32(S.KITTI_CX)) * zz).astype(f32) * invfx, ((k["y"] - f32(S.KITTI_CY)) * zz).astype(f32) * invfy, zz], 1).astype(f32)
    q_inv, ow = se3f_inverse(pose)
    R = quat_to_matrix_f32(q_inv)
    xw = np.stack([(R[r, 0] * pc[:, 0] + (R[r, 1] * pc[:, 1] + R[r, 2] * pc[:, 2]).astype(f32)).astype(f32) + ow[r] for r in range(3)], 1).astype(f32)
    return xw, ok


def oracle_chain(frames, sf, pose0, th=15.0):
    poses = [np.asarray(pose0, f32)]
    nms, nis = [0], [0]
    for t in range(1, len(frames)):
        last, cur = frames[t - 1], frames[t]
        lp = poses[-1]
        xw, ok = chain_unproject(last, lp)
        fv = oracle.FrameView(*frame_view_args(cur, sf))
        nm, match = oracle.search_by_projection_last(fv, lp, lp, ok.astype(np.uint8), xw, last["d"], last["k"]["octave"], last["k"]["angle"],
                                                     np.ones(len(ok), np.uint8), th)
        m = np.nonzero(match >= 0)[0]
        obs = np.stack([cur["k"]["x"][m], cur["k"]["y"][m], cur["ur"][m]], 1).astype(f32)
        s = sf[cur["k"]["octave"][m]]
        inv_s2 = (f32(1.0) / (s * s).astype(f32)).astype(f32)
        st = (cur["ur"][m] >= 0).astype(np.uint8)
        ni, pose, _ = oracle.pose_optimize(lp, xw[match[m]], obs, inv_s2, st, *CAM)
        poses.append(pose); nms.append(nm); nis.append(ni)
    return np.stack(poses), np.array(nms), np.array(nis)


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
