#!/usr/bin/env python3
"""Generate tests/golden/*.npz with the REAL OpenCV primitives (python cv2), i.e. the library the
reference links (CMakeLists.txt:33 `find_package(OpenCV 4.4)`), driven by tests/cv2_reference.py.
Run in the build container (cv2 4.13.0 there); the outputs are committed so that the oracle can be
pinned on boxes without cv2 and without /root/reference.

    python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cv2  # noqa: E402
import cv2_reference as R  # noqa: E402
import oracle  # noqa: E402
from orb_slam3_rgbl_b200 import synthetic as S  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def extractor_case(name, seed, W, H, nf):
    img = S.make_image(seed, W, H)
    ex = oracle.Extractor(nf)
    k, d, per, rois = R.extract_cv2(img, oracle.distribute_quadtree, nfeatures=nf, quota=ex.features_per_level)
    out = dict(seed=seed, W=W, H=H, nfeatures=nf, image_sha=sha(img), cv2_version=cv2.__version__,
               kps=k, desc=d)
    for l, (cand, sel) in enumerate(per):
        out[f"cand{l}"] = cand.astype(np.int16)
        out[f"level_sha{l}"] = sha(rois[l])
        out[f"blur_sha{l}"] = sha(cv2.GaussianBlur(rois[l].copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    np.savez_compressed(Path(__file__).parent / f"{name}.npz", **out)
    print(name, len(k), "keypoints")


def depth_case(name, seed, W, H, n_az):
    pts = S.make_pointcloud(seed, 64, n_az)
    P = S.lidar_projection_matrix()
    raw = R.project_cv2(pts, P, W, H)
    out = dict(seed=seed, W=W, H=H, n_az=n_az, pts_sha=sha(pts), P=P, raw_sha=sha(raw),
               raw_nonzero=np.argwhere(raw > 0).astype(np.int16), raw_vals=raw[raw > 0], cv2_version=cv2.__version__)
    for kind, ku, kv in (("diamond", 5, 5), ("rectangle", 5, 3), ("ellipse", 7, 5), ("cross", 3, 7), ("diamond", 9, 9)):
        m = S.structuring_element(kind, ku, kv)
        if kind != "diamond":
            cvm = cv2.getStructuringElement({"rectangle": cv2.MORPH_RECT, "cross": cv2.MORPH_CROSS, "ellipse": cv2.MORPH_ELLIPSE}[kind], (ku, kv))
            assert (cvm == m).all()
        proc = R.inverse_dilation_cv2(raw, m)
        out[f"proc_sha_{kind}_{ku}_{kv}"] = sha(proc)
        out[f"mask_{kind}_{ku}_{kv}"] = m
    # a few hundred probe pixels of the default (diamond 5) result
    proc = R.inverse_dilation_cv2(raw, S.structuring_element("diamond", 5))
    rng = np.random.default_rng(seed)
    yy = rng.integers(0, H, 400); xx = rng.integers(0, W, 400)
    out["probe_yx"] = np.stack([yy, xx], 1).astype(np.int16); out["probe_vals"] = proc[yy, xx]
    np.savez_compressed(Path(__file__).parent / f"{name}.npz", **out)
    print(name, int((raw > 0).sum()), "projected pixels")


def primitive_case():
    rng = np.random.default_rng(7)
    ys = rng.integers(-70000, 70000, 4000).astype(np.float32); xs = rng.integers(-70000, 70000, 4000).astype(np.float32)
    at = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    np.savez_compressed(Path(__file__).parent / "primitives.npz", atan_y=ys, atan_x=xs, atan_deg=at, cv2_version=cv2.__version__)


if __name__ == "__main__":
    extractor_case("extract_kitti_seed2", 2, S.KITTI_W, S.KITTI_H, 2000)
    extractor_case("extract_small_seed9", 9, 400, 300, 500)
    depth_case("depth_kitti_seed2", 2, S.KITTI_W, S.KITTI_H, 1875)
    primitive_case()
