"""ORACLE SUPPORT (test infrastructure, NOT product code): the tracking chain of rgbl_resident_track_begin2 composed from the
reference-pinned oracle functions (oracle.search_by_projection_last / is_in_frustum / search_by_projection_local / pose_optimize) with
the float32 glue between them written out in numpy exactly as the reference computes it (Frame::UnprojectStereo,
Frame::UpdatePoseMatrices, Sophus::SE3f::inverse, MapPoint::UpdateNormalAndDepth), each piece pinned against the reference's own code
in tests/test_oracle_tracking_ref.py.  Used by the GPU parity tests and by bench.py's CPU legs (cpu_baseline, --impl reference)."""
import numpy as np

import oracle

# KITTI 00-02 calibration of Examples/RGB-L/KITTI00-02.yaml (fx, fy, cx, cy, bf); the callers pass their own when it differs
f32 = np.float32


def frame_view_args(fr, sf, W, H, cam):
    return (fr["k"], fr["ur"], fr["d"], W, H, sf) + tuple(cam)




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


def se3f_mul(a, b):
    """Sophus::SE3f * SE3f (se3.hpp:304-308): (Ra Rb, ta + Ra tb).  SO3f * SO3f is the Hamilton product written out in so3.hpp:325-339,
    evaluated left to right in float32, and its result goes through the SO3f(quaternion) constructor, which normalises (so3.hpp:481-487,
    297-303: coeffs /= norm, Eigen's 4-term reduction (x2 + y2) + (z2 + w2)).  Poses are (qx, qy, qz, qw, tx, ty, tz) float32."""
    ax, ay, az, aw = (f32(v) for v in a[:4]); bx, by, bz, bw = (f32(v) for v in b[:4])
    w = f32(f32(f32(aw * bw) - f32(ax * bx)) - f32(ay * by)) - f32(az * bz)
    x = f32(f32(f32(aw * bx) + f32(ax * bw)) + f32(ay * bz)) - f32(az * by)
    y = f32(f32(f32(aw * by) + f32(ay * bw)) + f32(az * bx)) - f32(ax * bz)
    z = f32(f32(f32(aw * bz) + f32(az * bw)) + f32(ax * by)) - f32(ay * bx)
    q = np.array([x, y, z, w], f32)
    length = np.sqrt(f32(f32(q[0] * q[0]) + f32(q[1] * q[1])) + f32(f32(q[2] * q[2]) + f32(q[3] * q[3])))
    q = (q / f32(length)).astype(f32)
    t = (np.asarray(a[4:7], f32) + se3f_rotate(np.asarray(a[:4], f32), np.asarray(b[4:7], f32)[None, :])[0]).astype(f32)
    return np.concatenate([q, t]).astype(f32)


def predict_pose(prev_pose, last_pose):
    """Tracking::TrackWithMotionModel's initial pose mVelocity * mLastFrame.GetPose() (src/Tracking.cc:2904) with the constant-velocity
    model mVelocity = Tcw(last) * Tcw(prev)^-1 (src/Tracking.cc:2243-2245).  prev_pose None (no velocity yet, the reference then tracks
    against the reference key frame): the last pose itself."""
    last_pose = np.asarray(last_pose, f32)
    if prev_pose is None:
        return last_pose
    qi, ti = se3f_inverse(np.asarray(prev_pose, f32))
    velocity = se3f_mul(last_pose, np.concatenate([qi, ti]).astype(f32))
    return se3f_mul(velocity, last_pose)


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


def chain_unproject(fr, pose, cam):
    """Frame::UnprojectStereo (src/Frame.cc:1137-1150) for every keypoint: x3D = mRwc * x3Dc + mOw with mRwc / mOw from
    Frame::UpdatePoseMatrices (src/Frame.cc:562-569), float32, Eigen 3.3's product order r0*x + (r1*y + r2*z).
    Pinned to the reference's own code by tests/test_oracle_tracking_ref.py::test_unproject_stereo."""
    k, z = fr["k"], fr["depth"]
    ok = z > 0
    zz = np.where(ok, z, f32(1)).astype(f32)
    fx, fy, cx, cy = (f32(v) for v in cam[:4])
    invfx, invfy = f32(1.0) / fx, f32(1.0) / fy
    pc = np.stack([((k["x"] - cx) * zz).astype(f32) * invfx, ((k["y"] - cy) * zz).astype(f32) * invfy, zz], 1).astype(f32)
    q_inv, ow = se3f_inverse(pose)
    R = quat_to_matrix_f32(q_inv)
    xw = np.stack([(R[r, 0] * pc[:, 0] + (R[r, 1] * pc[:, 1] + R[r, 2] * pc[:, 2]).astype(f32)).astype(f32) + ow[r] for r in range(3)], 1).astype(f32)
    return xw, ok


def oracle_chain(frames, sf, pose0, W, H, cam, th=15.0):
    poses = [np.asarray(pose0, f32)]
    nms, nis = [0], [0]
    for t in range(1, len(frames)):
        last, cur = frames[t - 1], frames[t]
        lp = poses[-1]
        pred = predict_pose(poses[-2] if len(poses) >= 2 else None, lp)
        xw, ok = chain_unproject(last, lp, cam)
        fv = oracle.FrameView(*frame_view_args(cur, sf, W, H, cam))
        nm, match = oracle.search_by_projection_last(fv, pred, lp, ok.astype(np.uint8), xw, last["d"], last["k"]["octave"], last["k"]["angle"],
                                                     np.ones(len(ok), np.uint8), th)
        m = np.nonzero(match >= 0)[0]
        obs = np.stack([cur["k"]["x"][m], cur["k"]["y"][m], cur["ur"][m]], 1).astype(f32)
        s = sf[cur["k"]["octave"][m]]
        inv_s2 = (f32(1.0) / (s * s).astype(f32)).astype(f32)
        st = (cur["ur"][m] >= 0).astype(np.uint8)
        ni, pose, _ = oracle.pose_optimize(pred, xw[match[m]], obs, inv_s2, st, *cam)
        poses.append(pose); nms.append(nm); nis.append(ni)
    return np.stack(poses), np.array(nms), np.array(nis)


def _sum3(a, b, c):
    """Eigen 3.3's 3-term reduction a + (b + c), float32."""
    return (a + (b + c).astype(f32)).astype(f32)


def pose_matrices(pose):
    """Frame::UpdatePoseMatrices (src/Frame.cc:562-569) in float32: (mRcw, mtcw, mOw)."""
    R = quat_to_matrix_f32(np.asarray(pose[:4], f32))
    _, ow = se3f_inverse(pose)
    return R, np.asarray(pose[4:7], f32), ow


def local_points_of(fr, pose, sf, cam):
    """The LiDAR-depth keypoints of a frame as local map points: world position (Frame::UnprojectStereo with the frame's final pose),
    normal and scale-invariance distances of MapPoint::UpdateNormalAndDepth for one observation (src/MapPoint.cc:437-490):
    normal = PC / |PC|, mfMaxDistance = |PC| * scale[octave], mfMinDistance = mfMaxDistance / scale[nLevels - 1].  float32, Eigen order."""
    xw, ok = chain_unproject(fr, pose, cam)
    _, ow = se3f_inverse(pose)
    pc = (xw - ow[None, :]).astype(f32)
    dist = np.sqrt(_sum3(pc[:, 0] * pc[:, 0], pc[:, 1] * pc[:, 1], pc[:, 2] * pc[:, 2])).astype(f32)
    dist_safe = np.where(ok, dist, f32(1))
    normal = (pc / dist_safe[:, None]).astype(f32)
    mx = (dist * sf[fr["k"]["octave"]]).astype(f32)
    mn = (mx / sf[len(sf) - 1]).astype(f32)
    return dict(valid=ok, xw=xw, normal=normal, mn=mn, mx=mx, desc=fr["d"])


def oracle_chain2(frames, sf, pose0, W, H, cam, K=3, th_last=15.0, th_local=3.0, nn_ratio=0.8, state=None):
    """The chain of rgbl_resident_track_begin2 composed from the (reference-pinned) oracle functions, frame by frame:
    SearchByProjection(last) -> PoseOptimization -> discard outliers -> isInFrustum over the local map ring -> SearchByProjection(local)
    -> PoseOptimization on all map points.  `state` (returned as last element) carries the last frame, its pose and the ring into the
    next batch (continue_sequence).  -> (poses, n_matches, n_inliers, n_local_matches, n_inliers_first, state)"""
    if state is None:
        ring = [None] * max(K, 1); count = 0
        last, last_pose, prev_pose = frames[0], np.asarray(pose0, f32), None
        poses, out = [last_pose], [(0, 0, 0, 0)]
        todo = frames[1:]
    else:
        ring, count, last, last_pose, prev_pose = state["ring"], state["count"], state["last"], state["pose"], state.get("prev_pose")
        poses, out = [], []
        todo = frames
    for cur in todo:
        xw, ok = chain_unproject(last, last_pose, cam)
        pred = predict_pose(prev_pose, last_pose)                       # constant-velocity motion model
        fv = oracle.FrameView(*frame_view_args(cur, sf, W, H, cam))
        nm, match = oracle.search_by_projection_last(fv, pred, last_pose, ok.astype(np.uint8), xw, last["d"], last["k"]["octave"], last["k"]["angle"],
                                                     np.ones(len(ok), np.uint8), th_last)
        m = np.nonzero(match >= 0)[0]

        def edges(idx, pts):
            obs = np.stack([cur["k"]["x"][idx], cur["k"]["y"][idx], cur["ur"][idx]], 1).astype(f32)
            s = sf[cur["k"]["octave"][idx]]
            return pts, obs, (f32(1.0) / (s * s).astype(f32)).astype(f32), (cur["ur"][idx] >= 0).astype(np.uint8)

        ni1, pose1, outl = oracle.pose_optimize(pred, *edges(m, xw[match[m]]), *cam)
        if K == 0:
            pose2, ni2, nml = pose1, ni1, 0
        else:
            keep = m[outl == 0]                                         # outliers are discarded (src/Tracking.cc:2944-2966)
            cur_state = np.zeros(len(cur["k"]), np.uint8); cur_state[keep] = 1
            slots = [r for r in ring if r is not None]
            if slots:
                lp = {k: np.concatenate([r[k] for r in ring if r is not None]) for k in ("valid", "xw", "normal", "mn", "mx", "desc")}
                v = lp["valid"]
                R, tcw, ow = pose_matrices(pose1)
                tr = oracle.is_in_frustum(fv, R, tcw, ow, lp["xw"][v], lp["normal"][v], lp["mn"][v], lp["mx"][v], 0.5)
                nml, ml = oracle.search_by_projection_local(fv, tr, lp["desc"][v], np.ones(int(v.sum()), np.uint8), th_local, nn_ratio, False, 0.0, cur_state)
                lxw = lp["xw"][v]
            else:
                nml, ml, lxw = 0, np.full(len(cur["k"]), -1, np.int32), np.zeros((0, 3), f32)
            src_last = np.full(len(cur["k"]), -1, np.int64); src_last[keep] = match[keep]
            idx = np.nonzero((src_last >= 0) | (ml >= 0))[0]            # keypoint order = Optimizer::PoseOptimization's edge order
            pts = np.where((src_last[idx] >= 0)[:, None], xw[np.maximum(src_last[idx], 0)], lxw[np.maximum(ml[idx], 0)] if len(lxw) else xw[np.maximum(src_last[idx], 0)]).astype(f32)
            ni2, pose2, _ = oracle.pose_optimize(pose1, *edges(idx, pts), *cam)
            ring[count % K] = local_points_of(last, last_pose, sf, cam); count += 1      # the last frame's points join the local map
        poses.append(pose2); out.append((nm, ni2, nml, ni1))
        last, prev_pose, last_pose = cur, last_pose, pose2
    o = np.array(out, np.int64).reshape(-1, 4)
    return np.stack(poses), o[:, 0], o[:, 1], o[:, 2], o[:, 3], dict(ring=ring, count=count, last=last, pose=last_pose, prev_pose=prev_pose)


