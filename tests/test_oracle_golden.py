"""Pin the C++ oracle against the committed cv2-generated fixtures (tests/golden/make_golden.py) and,
when python-cv2 is importable, against cv2 live.  CPU only."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

import oracle
from orb_slam3_rgbl_b200 import synthetic as S

G = Path(__file__).resolve().parent / "golden"


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["extract_kitti_seed2", "extract_small_seed9"])
def test_extractor_against_cv2_fixture(name):
    g = np.load(G / f"{name}.npz")
    img = S.make_image(int(g["seed"]), int(g["W"]), int(g["H"]))
    assert sha(img) == str(g["image_sha"]), "synthetic generator drifted: regenerate the fixtures"
    ex = oracle.Extractor(int(g["nfeatures"]))
    k, d, mono = ex(img)
    for l in range(8):
        lv = ex.level_image(l)
        assert sha(lv) == str(g[f"level_sha{l}"]), f"pyramid level {l}"
        assert sha(oracle.gaussian_blur7(lv)) == str(g[f"blur_sha{l}"]), f"blur level {l}"
        assert (ex.level_candidates(l) == g[f"cand{l}"].astype(np.int32)).all(), f"FAST candidates level {l}"
    gk = g["kps"]
    assert len(k) == len(gk) == mono
    for f in k.dtype.names:
        assert (k[f] == gk[f]).all(), f
    assert (d == g["desc"]).all()


def test_depth_against_cv2_fixture():
    g = np.load(G / "depth_kitti_seed2.npz")
    W, H = int(g["W"]), int(g["H"])
    pts = S.make_pointcloud(int(g["seed"]), 64, int(g["n_az"]))
    assert sha(pts) == str(g["pts_sha"])
    raw = oracle.depth_project(pts, g["P"], W, H)
    assert sha(raw) == str(g["raw_sha"])
    nz = g["raw_nonzero"].astype(int)
    assert (raw[nz[:, 0], nz[:, 1]] == g["raw_vals"]).all()
    for key in g.files:
        if key.startswith("mask_"):
            kind = key[len("mask_"):]
            proc = oracle.depth_inverse_dilation(raw, g[key])
            assert sha(proc) == str(g[f"proc_sha_{kind}"]), kind
    proc = oracle.depth_inverse_dilation(raw, S.structuring_element("diamond", 5))
    yx = g["probe_yx"].astype(int)
    assert (proc[yx[:, 0], yx[:, 1]] == g["probe_vals"]).all()


def test_fast_atan2_fixture():
    g = np.load(G / "primitives.npz")
    got = np.array([oracle.fast_atan2(y, x) for y, x in zip(g["atan_y"], g["atan_x"])], np.float32)
    assert (got == g["atan_deg"]).all()


# ---- live cv2 (same library family the reference links); skipped where cv2 is absent ----
cv2 = pytest.importorskip("cv2", reason="python-cv2 not installed") if False else None
try:
    import cv2 as _cv2
    import cv2_reference as R
    HAVE = True
except Exception:  # pragma: no cover
    HAVE = False
needs_cv2 = pytest.mark.skipif(not HAVE, reason="python-cv2 not installed")


@needs_cv2
@pytest.mark.parametrize("shape", [(1241, 376, 1034, 313), (346, 105, 288, 88), (640, 480, 533, 400), (67, 70, 56, 58)])
def test_resize_live(shape):
    sw, sh, dw, dh = shape
    img = S.make_image(4, max(sw, 80), max(sh, 80))[:sh, :sw]
    assert (oracle.resize_linear(img, dw, dh) == _cv2.resize(np.ascontiguousarray(img), (dw, dh), interpolation=_cv2.INTER_LINEAR)).all()


@needs_cv2
def test_blur_and_fast_live():
    img = S.make_image(6, 500, 200)
    assert (oracle.gaussian_blur7(img) == _cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=_cv2.BORDER_REFLECT_101)).all()
    rng = np.random.default_rng(0)
    for th in (7, 12, 20):
        det = _cv2.FastFeatureDetector_create(th, True, _cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        for _ in range(40):
            x, y = int(rng.integers(0, 440)), int(rng.integers(0, 150)); w, h = int(rng.integers(7, 50)), int(rng.integers(7, 50))
            win = img[y:y + h, x:x + w]
            ref = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in det.detect(win)], np.int32).reshape(-1, 3)
            got = oracle.fast_window(win, th)
            assert got.shape == ref.shape and (got == ref).all()


@needs_cv2
def test_full_extractor_live():
    img = S.make_image(13, 800, 300)
    ex = oracle.Extractor(1200)
    k, d, _ = ex(img)
    k2, d2, per, rois = R.extract_cv2(img, oracle.distribute_quadtree, nfeatures=1200, quota=ex.features_per_level)
    assert len(k) == len(k2) and all((k[f] == k2[f]).all() for f in k.dtype.names) and (d == d2).all()


@needs_cv2
def test_depth_live():
    pts = S.make_pointcloud(5); P = S.lidar_projection_matrix()
    raw = oracle.depth_project(pts, P, S.KITTI_W, S.KITTI_H)
    assert (raw == R.project_cv2(pts, P, S.KITTI_W, S.KITTI_H)).all()
    m = S.structuring_element("diamond", 7)
    assert (oracle.depth_inverse_dilation(raw, m) == R.inverse_dilation_cv2(raw, m)).all()


def test_glibc_sincos_is_what_the_descriptor_sees():
    """The oracle's descriptor uses libm cosf/sinf (canonical semantics, SURVEY hard part 4)."""
    img = S.make_image(1, 200, 200)
    blur = oracle.gaussian_blur7(img)
    kp = np.zeros(1, oracle.KP_DTYPE); kp["x"], kp["y"], kp["angle"] = 100, 100, 37.25
    d0 = oracle.descriptor(kp[0], blur)
    if HAVE:
        d1 = R.descriptor_np(blur, 100, 100, 37.25, R.load_pattern())
        assert (d0 == d1).all()
    assert d0.shape == (32,)


@needs_cv2
def test_other_upsamplers_live():
    """AverageFiltering (filter2D with FMA accumulation) and NearestNeighborPixel (fixed-point chamfer distanceTransform)."""
    pts = S.make_pointcloud(3); P = S.lidar_projection_matrix()
    raw = oracle.depth_project(pts, P, S.KITTI_W, S.KITTI_H)
    for k in (3, 5):
        a = oracle.depth_average_filter(raw, k)
        F = _cv2.filter2D(raw, -1, np.ones((k, k), np.float32) / (k * k), anchor=(-1, -1), delta=0, borderType=_cv2.BORDER_DEFAULT)
        _, B = _cv2.threshold(raw, 0, 1, 0)
        Cn = _cv2.filter2D(B, -1, np.ones((k, k), np.uint16), anchor=(-1, -1), delta=0, borderType=_cv2.BORDER_DEFAULT)
        with np.errstate(all="ignore"):
            ref = _cv2.multiply(F, _cv2.divide(float(k * k), Cn))
        m = ~(np.isnan(a) & np.isnan(ref))
        assert (np.isnan(a) == np.isnan(ref)).all() and (a[m] == ref[m]).all()
    R = 7
    rng = np.random.default_rng(0); n = 800
    k = np.zeros(n, oracle.KP_DTYPE); k["x"] = rng.uniform(19, S.KITTI_W - 19, n).astype(np.float32); k["y"] = rng.uniform(19, S.KITTI_H - 19, n).astype(np.float32)
    pad = _cv2.copyMakeBorder(raw, R, R, R, R, _cv2.BORDER_CONSTANT, value=0)
    D8 = np.clip(np.rint(raw), 0, 255).astype(np.uint8)
    _, DM = _cv2.threshold(D8, 0, 1, _cv2.THRESH_BINARY_INV)
    dist, _ = _cv2.distanceTransformWithLabels(DM, _cv2.DIST_L2, 5)
    dref = np.full(n, -1, np.float32)
    for i in range(n):
        u, v = k["x"][i], k["y"][i]
        sr = int(dist[int(v), int(u)]); d = 0.0
        if 0 <= sr < R:
            sr += 1
            x0 = int(np.float32(u) + np.float32(R) - np.float32(sr)); y0 = int(np.float32(v) + np.float32(R) - np.float32(sr))
            d = float(pad[y0:y0 + 2 * sr, x0:x0 + 2 * sr].max())
        if d > 0:
            dref[i] = d
    d, _u = oracle.depth_nearest_neighbor_pixel(raw, k, k, 100.0, float(R))
    assert (d == dref).all()


@needs_cv2
def test_descriptor_distance_is_cv2_norm_hamming():
    """ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2058-2074) is the bit count of the XOR = cv2.norm(a, b, NORM_HAMMING), which is
    also what DBoW2's FORB::distance (FORB.cpp:81-101) computes; the matcher / BoW / mapping oracles all rest on it."""
    import cv2 as _cv2
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (300, 32), dtype=np.uint8); b = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    b[:20] = a[:20]; b[20:40] = ~a[20:40]
    for i in range(len(a)):
        assert oracle.descriptor_distance(a[i], b[i]) == int(_cv2.norm(a[i], b[i], _cv2.NORM_HAMMING))
    # the nearest descriptor found by cv2's brute-force matcher has the minimum oracle distance
    bf = _cv2.BFMatcher(_cv2.NORM_HAMMING)
    for m in bf.match(a[:50], b):
        assert m.distance == min(oracle.descriptor_distance(a[m.queryIdx], x) for x in b)


@needs_cv2
def test_pose_oracle_agrees_with_cv2_solvepnp_on_an_inlier_only_problem():
    """Independent solver check for the PoseOptimization restatement: on a monocular, outlier-free problem with unit information the
    last (non-robust) round minimises the plain reprojection error, which is what cv2.solvePnP(SOLVEPNP_ITERATIVE) minimises."""
    import cv2 as _cv2
    import tracking_data as TD
    fx, fy, cx, cy, bf = TD.CAM
    for seed in (1, 2):
        pr = TD.pose_problem(seed, n=400, outlier_frac=0.0, stereo_frac=0.0)
        inv_s2 = np.ones(len(pr["xw"]), np.float32)
        n_in, pose, out = oracle.pose_optimize(pr["pose0"], pr["xw"], pr["obs"], inv_s2, pr["stereo"], *TD.CAM)
        keep = out == 0
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
        ok, rvec, tvec = _cv2.solvePnP(pr["xw"][keep].astype(np.float64), pr["obs"][keep, :2].astype(np.float64), K, None,
                                       rvec=np.zeros((3, 1)), tvec=np.zeros((3, 1)), useExtrinsicGuess=True, flags=_cv2.SOLVEPNP_ITERATIVE)
        assert ok
        Rcv, _ = _cv2.Rodrigues(rvec)
        x, y, z, w = pose[:4].astype(np.float64)
        Ror = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(Ror - Rcv).max() < 2e-4, np.abs(Ror - Rcv).max()
        assert np.abs(pose[4:].astype(np.float64) - tvec.ravel()).max() < 5e-3, np.abs(pose[4:] - tvec.ravel()).max()
        assert keep.mean() > 0.8          # the generator scales the pixel noise with the octave; with unit information the noisiest points are cut


def test_oracle_refuses_the_reference_division_by_zero_geometry():
    """DistributeOctTree divides by nIni = round(window width / window height) (src/ORBextractor.cc:559-561); for portrait
    pyramid levels that is 0 and the reference has undefined behaviour.  The oracle reports it instead of restating a crash."""
    import pytest
    from orb_slam3_rgbl_b200 import synthetic as S
    with pytest.raises(RuntimeError):
        oracle.Extractor(500)(S.make_image(1, 300, 560))


# ---- the extractor oracle against the reference's own ORBextractor.cc (oracle/_ref/libref_orbextractor.so) -----------------
@pytest.mark.parametrize("seed,size,prm", [(0, (900, 300), (1500, 1.2, 8, 12, 7)), (1, (1241, 376), (2000, 1.2, 8, 20, 7)),
                                          (2, (640, 480), (1000, 1.2, 8, 12, 7)), (3, (752, 480), (1200, 1.1, 6, 15, 5)),
                                          (4, (500, 400), (300, 1.6, 4, 12, 7)), (5, (1241, 376), (5000, 1.2, 8, 7, 3))])
def test_extractor_oracle_equals_the_reference_orbextractor(seed, size, prm):
    """oracle.Extractor vs ORBextractor::operator() compiled unmodified from /root/reference/src/ORBextractor.cc (tables, pyramid
    views, cell grid and fallback, DistributeOctTree with the real std::list / std::sort, IC_Angle, steered BRIEF, output order);
    the OpenCV primitives under it are the oracle's restatements, pinned against python-cv2 by the tests above."""
    import oracle as O
    from orb_slam3_rgbl_b200 import synthetic as S
    if O.ref_orbextractor() is None:
        pytest.skip("oracle/_ref/libref_orbextractor.so is not built and /root/reference is not available")
    w, h = size
    img = S.make_image(700 + seed, w, h)
    nf, sf, nl, ini, mn = prm
    for lap in ((0, 0), (w // 3, w // 2)):
        rk, rd, rmono = O.ref_orb_extract(img, nf, sf, nl, ini, mn, lap)
        ok, od, omono = O.Extractor(nf, sf, nl, ini, mn)(img, lap)
        assert len(ok) == len(rk) > 100 and omono == rmono
        for f in ok.dtype.names:
            assert (ok[f].view(np.uint32) == rk[f].view(np.uint32)).all(), f
        assert (od == rd).all()


def test_orb_tables_equal_the_reference_constructor():
    import oracle as O
    L = O.ref_orbextractor()
    if L is None:
        pytest.skip("reference build not available")
    for nf, sf, nl in ((2000, 1.2, 8), (1000, 1.1, 6), (500, 2.0, 4), (1250, 1.44, 5)):
        a, b, c, d = (np.empty(nl, np.float32) for _ in range(4))
        assert L.ref_orb_tables(nf, sf, nl, O._p(a), O._p(b), O._p(c), O._p(d)) == nl
        ex = O.Extractor(nf, sf, nl)
        assert (ex.scale_factors.view(np.uint32) == a.view(np.uint32)).all() and (ex.inv_scale_factors.view(np.uint32) == b.view(np.uint32)).all()


# ---- the DepthModule oracle against the reference's own DepthModule.cc (oracle/_ref/libref_depthmodule.so) ------------------
@pytest.mark.parametrize("kind,k,seed", [("Diamond", 5, 0), ("Rectangle", 3, 1), ("Cross", 7, 2), ("Ellipse", 5, 3), ("Diamond", 9, 4)])
def test_depth_oracle_equals_the_reference_depthmodule(tmp_path, kind, k, seed):
    """The reference parses the settings itself (projection matrix K [R|t], limits, method, structuring element), projects the
    cloud, runs Upsample_InverseDilation and GetFeatureDepthFromDepthMap; oracle.depth_from_pcd gets the reference's 12 projection
    floats and must return the same RawDepthMap, ProcessedDepthMap, mvDepth and mvuRight bit for bit."""
    import oracle as O
    from orb_slam3_rgbl_b200 import synthetic as S
    if O.ref_depthmodule() is None:
        pytest.skip("oracle/_ref/libref_depthmodule.so is not built and /root/reference is not available")
    K, Tr = S.camera_matrix(), S.KITTI_TR
    lines = [f"Camera.fx {float(K[0, 0])!r}", f"Camera.fy {float(K[1, 1])!r}", f"Camera.cx {float(K[0, 2])!r}", f"Camera.cy {float(K[1, 2])!r}",
             f"Camera.bf {float(S.KITTI_BF)!r}", "LiDAR.min_dist 5.0", "LiDAR.max_dist 200.0", "LiDAR.Method InverseDilation",
             f"LiDAR.MethodInverseDilation.KernelType {kind}", f"LiDAR.MethodInverseDilation.KernelSize_u {k}.0",
             f"LiDAR.MethodInverseDilation.KernelSize_v {k}.0"]
    lines += [f"LiDAR.Tr{r + 1}{c + 1} {float(Tr[r, c])!r}" for r in range(3) for c in range(4)]
    path = tmp_path / "settings.txt"; path.write_text("\n".join(lines) + "\n")
    W, H = S.KITTI_W, S.KITTI_H
    pts = S.make_pointcloud(10 + seed, n_azimuth=700)
    rng = np.random.default_rng(seed)
    kp = np.zeros(600, O.KP_DTYPE); kp["x"] = rng.integers(0, W, 600); kp["y"] = rng.integers(0, H, 600)
    ku = kp.copy(); ku["x"] += rng.normal(0, 0.3, 600).astype(np.float32)
    P, raw, proc, d, u = O.ref_depth_from_pcd(path, pts, W, H, np.stack([kp["x"], kp["y"]], 1), np.stack([ku["x"], ku["y"]], 1))
    # ParseRGBLParameters stores K and Tr as float32 and multiplies them with cv::gemm (double accumulation, one rounding)
    K4 = np.zeros((3, 4), np.float32); K4[:, :3] = K.astype(np.float32)
    T4 = np.vstack([Tr.astype(np.float32), np.array([[0, 0, 0, 1]], np.float32)])
    assert (P.view(np.uint32) == (K4.astype(np.float64) @ T4.astype(np.float64)).astype(np.float32).view(np.uint32)).all()
    od, ou, oraw, oproc = O.depth_from_pcd(pts, P, W, H, S.structuring_element(kind.lower(), k), S.KITTI_BF, kp, ku)
    assert (raw > 0).sum() > 3000
    assert (oraw.view(np.uint32) == raw.view(np.uint32)).all()
    assert (oproc.view(np.uint32) == proc.view(np.uint32)).all()
    assert (od.view(np.uint32) == d.view(np.uint32)).all() and (ou.view(np.uint32) == u.view(np.uint32)).all()
    assert (d > 0).sum() > 50
