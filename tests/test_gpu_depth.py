"""GPU parity: CUDA DepthModule path vs the CPU oracle (bit-exact float32)."""
import numpy as np
import pytest

import oracle
from orb_slam3_rgbl_b200 import frontend as F
from orb_slam3_rgbl_b200 import synthetic as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=3, max_points=140000)
    yield c
    c.close()


def _kps(seed, n=1500):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, oracle.KP_DTYPE)
    k["x"] = rng.uniform(19, S.KITTI_W - 19, n).astype(np.float32)
    k["y"] = rng.uniform(19, S.KITTI_H - 19, n).astype(np.float32)
    return k


@pytest.mark.parametrize("kind,ku,kv", [("Diamond", 5, 5), ("Diamond", 9, 9), ("Rectangle", 5, 3), ("Cross", 7, 5), ("Ellipse", 7, 5)])
def test_depth_from_pcd(ctx, kind, ku, kv):
    pts = S.make_pointcloud(4); P = S.lidar_projection_matrix(); k = _kps(1)
    dm = F.DepthModule(ctx, P, S.KITTI_BF, "InverseDilation", 5.0, 200.0, kind, ku, kv)
    dm.CalculateDepthFromPcd(k, k, pts, S.KITTI_W, S.KITTI_H)
    mask = S.structuring_element(kind, ku, kv)
    d, u, raw, proc = oracle.depth_from_pcd(pts, P, S.KITTI_W, S.KITTI_H, mask, S.KITTI_BF, k, k)
    assert (dm.RawDepthMap == raw).all(), f"raw mismatches {(dm.RawDepthMap != raw).sum()}"
    assert (dm.ProcessedDepthMap == proc).all(), f"processed mismatches {(dm.ProcessedDepthMap != proc).sum()}"
    assert (dm.mvDepth == d).all() and (dm.mvuRight == u).all()
    assert (d > 0).sum() > 100


def test_last_writer_wins_and_stale_frames(ctx):
    """Two different clouds back to back on the same context: no leakage from the previous frame, and
    duplicate pixels resolved in favour of the later point (src/DepthModule.cc:123-137)."""
    P = S.lidar_projection_matrix(); k = _kps(2, 10)
    dm = F.DepthModule(ctx, P, S.KITTI_BF)
    a = S.make_pointcloud(8)
    b = np.ascontiguousarray(np.concatenate([a[:, :50000], a[:, :50000] * np.array([[1.01], [1.0], [1.0], [1.0]], np.float32)], axis=1))
    for pts in (a, b, a[:, :100]):
        dm.CalculateDepthFromPcd(k, k, pts, S.KITTI_W, S.KITTI_H)
        raw = oracle.depth_project(pts, P, S.KITTI_W, S.KITTI_H)
        assert (dm.RawDepthMap == raw).all()


def test_empty_cloud(ctx):
    P = S.lidar_projection_matrix(); k = _kps(3, 50)
    dm = F.DepthModule(ctx, P, S.KITTI_BF)
    dm.CalculateDepthFromPcd(k, k, np.zeros((4, 0), np.float32), S.KITTI_W, S.KITTI_H)
    assert (dm.mvDepth == -1).all() and (dm.mvuRight == -1).all() and (dm.ProcessedDepthMap == 0).all()


def test_frame_rgbl_batch(ctx):
    seeds = (31, 32, 33)
    imgs = [S.make_image(s) for s in seeds]; pcs = [S.make_pointcloud(s) for s in seeds]
    pcs[1] = np.ascontiguousarray(pcs[1][:, :90001])            # ragged batch
    P = S.lidar_projection_matrix()
    prm = F.make_depth_params(bf=S.KITTI_BF)
    outs = F.frame_rgbl_batch(ctx, imgs, pcs, P, prm)
    mask = S.structuring_element("diamond", 5)
    ref = oracle.Extractor(2000)
    for img, pc, (k, d, dep, ur) in zip(imgs, pcs, outs):
        rk, rd, _ = ref(img)
        assert len(k) == len(rk) and all((k[f] == rk[f]).all() for f in k.dtype.names) and (d == rd).all()
        rdep, rur, _, _ = oracle.depth_from_pcd(pc, P, S.KITTI_W, S.KITTI_H, mask, S.KITTI_BF, rk, rk)
        assert (dep == rdep).all() and (ur == rur).all()


def test_raw_kitti_records_equal_host_relayout(ctx):
    """rgbl_resident_upload_kitti: the .bin records (x, y, z, reflectance) de-interleaved on the device must give exactly what the
    reference's host loop (LoadPointcloudBinaryMat: rows x, y, z, 1) gives; ragged and empty clouds included."""
    seeds = (41, 42, 43)
    imgs = [S.make_image(s) for s in seeds]; pcs = [S.make_pointcloud(s) for s in seeds]
    pcs[1] = np.ascontiguousarray(pcs[1][:, :77777]); pcs[2] = np.ascontiguousarray(pcs[2][:, :0])
    rng = np.random.default_rng(0)
    raw = [np.ascontiguousarray(np.column_stack([pc[0], pc[1], pc[2], rng.random(pc.shape[1], dtype=np.float32)])) for pc in pcs]
    prm = F.make_depth_params(bf=S.KITTI_BF)
    b = F.RgblBatch(ctx, imgs, [pc if pc.shape[1] else np.zeros((4, 1), np.float32) for pc in pcs], S.lidar_projection_matrix(), prm, pinned=False)
    b.npts[:] = [pc.shape[1] for pc in pcs]
    b.upload(); b.process_resident(); ref = b.download()
    ref = [tuple(a.copy() for a in fr) for fr in ref]
    b.upload_kitti(raw); b.process_resident(); got = b.download()
    for (rk, rd, rdep, rur), (k, d, dep, ur) in zip(ref, got):
        assert np.array_equal(k, rk) and np.array_equal(d, rd) and np.array_equal(dep, rdep) and np.array_equal(ur, rur)
    assert (got[0][2] > 0).sum() > 100 and (got[2][2] > 0).sum() == 0        # the empty cloud gives no depth


@pytest.mark.parametrize("k", [5, 3, 7])
def test_average_filtering(ctx, k):
    """LiDAR.Method AverageFiltering (src/DepthModule.cc:200-228): bit-exact incl. the NaN pattern of empty windows."""
    pts = S.make_pointcloud(6); P = S.lidar_projection_matrix(); kp = _kps(7)
    dm = F.DepthModule(ctx, P, S.KITTI_BF, "AverageFiltering", avg_kernel=k)
    dm.CalculateDepthFromPcd(kp, kp, pts, S.KITTI_W, S.KITTI_H)
    raw = oracle.depth_project(pts, P, S.KITTI_W, S.KITTI_H)
    with np.errstate(all="ignore"):
        proc = oracle.depth_average_filter(raw, k)
    d, u = oracle.depth_gather(proc, kp, kp, S.KITTI_BF)
    assert (dm.RawDepthMap == raw).all()
    assert (np.isnan(dm.ProcessedDepthMap) == np.isnan(proc)).all()
    m = ~np.isnan(proc)
    assert (dm.ProcessedDepthMap[m] == proc[m]).all()
    assert (dm.mvDepth == d).all() and (dm.mvuRight == u).all() and (d > 0).sum() > 100


@pytest.mark.parametrize("R", [7.0, 3.0, 12.0])
def test_nearest_neighbor_pixel(ctx, R):
    """LiDAR.Method NearestNeighborPixel (src/DepthModule.cc:145-198), incl. the fixed-point chamfer distance transform."""
    pts = S.make_pointcloud(8); P = S.lidar_projection_matrix(); kp = _kps(9, 1800)
    dm = F.DepthModule(ctx, P, S.KITTI_BF, "NearestNeighborPixel", nn_radius=R)
    dm.CalculateDepthFromPcd(kp, kp, pts, S.KITTI_W, S.KITTI_H)
    raw = oracle.depth_project(pts, P, S.KITTI_W, S.KITTI_H)
    d, u = oracle.depth_nearest_neighbor_pixel(raw, kp, kp, S.KITTI_BF, R)
    assert (dm.mvDepth == d).all() and (dm.mvuRight == u).all()
    assert (d > 0).sum() > 200


def test_method_none_projects_only(ctx):
    pts = S.make_pointcloud(8); P = S.lidar_projection_matrix(); kp = _kps(9, 100)
    dm = F.DepthModule(ctx, P, S.KITTI_BF, "None")
    dm.CalculateDepthFromPcd(kp, kp, pts, S.KITTI_W, S.KITTI_H)
    assert (dm.RawDepthMap == oracle.depth_project(pts, P, S.KITTI_W, S.KITTI_H)).all()
    assert (dm.mvDepth == -1).all()
