"""The reference-side binding (shim/): the four drop-in classes compiled against the reference's own headers and the reference's
code around the replaced functions (tests/shim/shim_driver.cpp, tests/shim/Makefile), linked with librgbl_b200.so and driven by
Frame-constructor / Tracking-shaped caller code.
  * CPU: it compiles and links; without a CUDA device it fails loudly (no CPU fallback).
  * GPU: its results equal the oracle's (= the reference's, see test_oracle_tracking_ref.py) for frame construction, both
    SearchByProjection calls and both PoseOptimization calls of a tracked frame."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
import tracking_data as TD
from orb_slam3_rgbl_b200 import synthetic as S

HERE = Path(__file__).resolve().parent
DRIVER = HERE / "shim" / "build" / "shim_driver"


def build_driver():
    if Path("/root/reference/src/ORBmatcher.cc").exists():
        from orb_slam3_rgbl_b200 import _lib
        _lib.build()
        subprocess.run(["make", "-C", str(HERE / "shim")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    if not DRIVER.exists():
        pytest.skip("tests/shim/build/shim_driver is not built and /root/reference is not available")


def write_settings(path, kind="Diamond", k=5):
    K, Tr = S.camera_matrix(), S.KITTI_TR
    lines = [f"Camera.fx {float(K[0, 0])!r}", f"Camera.fy {float(K[1, 1])!r}", f"Camera.cx {float(K[0, 2])!r}", f"Camera.cy {float(K[1, 2])!r}",
             f"Camera.bf {float(S.KITTI_BF)!r}", "LiDAR.min_dist 5.0", "LiDAR.max_dist 200.0", "LiDAR.Method InverseDilation",
             f"LiDAR.MethodInverseDilation.KernelType {kind}", f"LiDAR.MethodInverseDilation.KernelSize_u {k}.0",
             f"LiDAR.MethodInverseDilation.KernelSize_v {k}.0"]
    lines += [f"LiDAR.Tr{r + 1}{c + 1} {float(Tr[r, c])!r}" for r in range(3) for c in range(4)]
    path.write_text("\n".join(lines) + "\n")


def write_input(path, seq):
    im0, im1, pc0, pc1 = seq.image(0), seq.image(1), seq.cloud(0), seq.cloud(1)
    with open(path, "wb") as f:
        f.write(struct.pack("4i", seq.W, seq.H, pc0.shape[1], pc1.shape[1]))
        f.write(np.asarray(TD.CAM, np.float32).tobytes()); f.write(np.asarray(seq.pose(0), np.float32).tobytes())
        f.write(im0.tobytes()); f.write(im1.tobytes()); f.write(np.ascontiguousarray(pc0, np.float32).tobytes()); f.write(np.ascontiguousarray(pc1, np.float32).tobytes())


def test_binding_compiles_links_and_fails_loudly_without_a_gpu(tmp_path, have_gpu):
    build_driver()
    if have_gpu:
        pytest.skip("a CUDA device is present: see the GPU test")
    seq = S.PlaneSequence(41, 3)
    write_settings(tmp_path / "settings.txt"); write_input(tmp_path / "in.bin", seq)
    r = subprocess.run([str(DRIVER), str(tmp_path / "settings.txt"), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr[-500:])
    assert "librgbl_b200" in r.stderr and not (tmp_path / "out.bin").exists()


@pytest.mark.gpu
def test_binding_equals_the_oracle_on_a_tracked_frame(tmp_path):
    build_driver()
    seq = S.PlaneSequence(41, 3)
    write_settings(tmp_path / "settings.txt"); write_input(tmp_path / "in.bin", seq)
    r = subprocess.run([str(DRIVER), str(tmp_path / "settings.txt"), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    raw = (tmp_path / "out.bin").read_bytes()
    hdr = np.frombuffer(raw, np.int32, 10); off = 40
    nL, nC, nmatches, inl1, nlocal, inl2, hd, pyr0_w, pyr7_h, pyr_probe = (int(v) for v in hdr)

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype, n, off); off += a.nbytes
        return a

    got = []
    for n in (nL, nC):
        got.append(dict(k=take(oracle.KP_DTYPE, n), d=take(np.uint8, n * 32).reshape(n, 32), depth=take(np.float32, n), ur=take(np.float32, n)))
    match1, outlier1, pose1 = take(np.int32, nC), take(np.uint8, nC), take(np.float32, 7)
    match2, pose2, outlier2 = take(np.int32, nC), take(np.float32, 7), take(np.uint8, nC)
    processed = take(np.float32, seq.W * seq.H).reshape(seq.H, seq.W)

    # ---- the same through the oracle ----
    # DepthModule::ParseRGBLParameters stores K and Tr as float32 and multiplies them with cv::gemm (double accumulation, one rounding)
    K4 = np.zeros((3, 4), np.float32); K4[:, :3] = S.camera_matrix()
    T4 = np.vstack([S.KITTI_TR.astype(np.float32), np.array([[0, 0, 0, 1]], np.float32)])
    seq.P = (K4.astype(np.float64) @ T4.astype(np.float64)).astype(np.float32)
    frames, sf = TD.extract_frames(seq, [0, 1])
    for g, o in zip(got, frames):
        assert len(g["k"]) == len(o["k"]) and all((g["k"][f] == o["k"][f]).all() for f in o["k"].dtype.names)
        assert (g["d"] == o["d"]).all() and (g["depth"] == o["depth"]).all() and (g["ur"] == o["ur"]).all()
    last, cur = frames
    assert hd == oracle.descriptor_distance(last["d"][0], cur["d"][0])
    ex = oracle.Extractor(2000); ex(seq.image(0))               # the driver probes the pyramid right after the first frame
    assert pyr0_w == seq.W and pyr7_h == ex.level_image(7).shape[0] and pyr_probe == int(ex.level_image(1)[10, 10])
    p0 = seq.pose(0)
    xw, ok = TD.chain_unproject(last, p0)
    fv = oracle.FrameView(*TD.frame_view_args(cur, sf))
    nm, m = oracle.search_by_projection_last(fv, p0, p0, ok.astype(np.uint8), xw, last["d"], last["k"]["octave"], last["k"]["angle"], np.ones(len(ok), np.uint8), 15.0)
    assert nm == nmatches and (np.where(m >= 0, m, -1) == match1).all()
    idx = np.nonzero(m >= 0)[0]

    def edges(ix, pts):
        obs = np.stack([cur["k"]["x"][ix], cur["k"]["y"][ix], cur["ur"][ix]], 1).astype(np.float32)
        s = sf[cur["k"]["octave"][ix]]
        return pts, obs, (np.float32(1.0) / (s * s).astype(np.float32)).astype(np.float32), (cur["ur"][ix] >= 0).astype(np.uint8)

    ni1, rpose1, rout1 = oracle.pose_optimize(p0, *edges(idx, xw[m[idx]]), *TD.CAM)
    assert ni1 == inl1 and (rout1 == outlier1[idx]).all() and np.abs(rpose1 - pose1).max() < 2e-5
    # local map = the first frame's points in index order; the matched ones are not searched again
    keep = idx[rout1 == 0]
    state = np.zeros(len(cur["k"]), np.uint8); state[keep] = 1
    lp = TD.local_points_of(last, p0, sf)
    v = lp["valid"]
    vi = np.nonzero(v)[0]
    R, tcw, ow = TD.pose_matrices(pose1)                       # the binding's pose after the first optimisation
    tr = oracle.is_in_frustum(fv, R, tcw, ow, lp["xw"][v], lp["normal"][v], lp["mn"][v], lp["mx"][v], 0.5)
    already = np.isin(vi, m[keep])
    tr["in_view"] = np.where(already, 0, tr["in_view"]).astype(np.uint8)
    nl, ml = oracle.search_by_projection_local(fv, tr, lp["desc"][v], np.ones(len(vi), np.uint8), 3.0, 0.8, False, 50.0, state)
    assert nl == nlocal
    exp2 = np.full(len(cur["k"]), -1, np.int64); exp2[keep] = m[keep]
    exp2 = np.where(ml >= 0, vi[np.maximum(ml, 0)], exp2)
    assert (exp2 == match2).all()
    ix2 = np.nonzero(exp2 >= 0)[0]
    ni2, rpose2, rout2 = oracle.pose_optimize(pose1, *edges(ix2, xw[exp2[ix2]]), *TD.CAM)
    assert ni2 == inl2 and (rout2 == outlier2[ix2]).all() and np.abs(rpose2 - pose2).max() < 2e-5
    assert abs(pose2[4] - seq.pose(1)[4]) < 0.02 and nlocal > 50
    _, _, _, oproc = oracle.depth_from_pcd(seq.cloud(1), seq.P, seq.W, seq.H, S.structuring_element("diamond", 5), S.KITTI_BF, cur["k"], cur["k"])
    assert (oproc == processed).all()
