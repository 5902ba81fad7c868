"""Python + cv2 restatement of the reference's orchestration, using the REAL OpenCV primitives
(cv2.resize / copyMakeBorder / FAST / GaussianBlur / fastAtan2 / gemm / dilate / threshold) exactly
where the reference calls them.  It is used (a) to pin the C++ oracle when cv2 is importable and
(b) by tests/golden/make_golden.py to generate the committed fixtures.

Reference: src/ORBextractor.cc:781-896,1086-1195 and src/DepthModule.cc:106-139,230-274.
The quad-tree has no OpenCV counterpart; it is injected as a callable.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

try:
    import cv2
    cv2.setNumThreads(1)
    HAVE_CV2 = True
except Exception:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False

_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float; _libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float; _libm.sinf.argtypes = [ctypes.c_float]

EDGE = 19
HALF_PATCH = 15
f32 = np.float32


def cv_round(v) -> int:
    return int(np.rint(f32(v)))


def scale_tables(nlevels=8, sf=1.2):
    sc = [f32(1.0)]
    for _ in range(1, nlevels):
        sc.append(f32(sc[-1] * f32(sf)))
    inv = [f32(f32(1.0) / s) for s in sc]
    return sc, inv


def pyramid_cv2(img: np.ndarray, nlevels=8, sf=1.2):
    """ORBextractor::ComputePyramid with padded planes; returns (padded planes, ROI views)."""
    sc, inv = scale_tables(nlevels, sf)
    H, W = img.shape
    padded, rois = [], []
    for l in range(nlevels):
        w, h = cv_round(f32(W) * inv[l]), cv_round(f32(H) * inv[l])
        if l == 0:
            temp = cv2.copyMakeBorder(img, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)
        else:
            lvl = cv2.resize(rois[l - 1], (w, h), interpolation=cv2.INTER_LINEAR)
            temp = cv2.copyMakeBorder(lvl, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)
        padded.append(temp)
        rois.append(temp[EDGE:EDGE + h, EDGE:EDGE + w])
    return padded, rois


def cell_candidates_cv2(roi: np.ndarray, ini_th=12, min_th=7):
    """Per-cell cv2 FAST with fallback; returns int32 [n,3] (x,y,score) relative to minBorder (16,16)."""
    h, w = roi.shape
    minBX = minBY = EDGE - 3
    maxBX, maxBY = w - EDGE + 3, h - EDGE + 3
    width, height = f32(maxBX - minBX), f32(maxBY - minBY)
    nCols, nRows = int(width / f32(35)), int(height / f32(35))
    wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
    det = {t: cv2.FastFeatureDetector_create(t, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16) for t in (ini_th, min_th)}
    out = []
    for i in range(nRows):
        iniY = minBY + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minBX + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            win = roi[iniY:maxY, iniX:maxX]          # strided view, like rowRange/colRange
            kps = det[ini_th].detect(win)
            if len(kps) == 0:
                kps = det[min_th].detect(win)
            for k in kps:
                out.append((int(k.pt[0]) + j * wCell, int(k.pt[1]) + i * hCell, int(k.response)))
    return np.array(out, np.int32).reshape(-1, 3)


def umax_table():
    um = [0] * 16
    vmax = int(math.floor(HALF_PATCH * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(HALF_PATCH * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        um[v] = int(np.rint(math.sqrt(HALF_PATCH * HALF_PATCH - v * v)))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while um[v0] == um[v0 + 1]:
            v0 += 1
        um[v] = v0
        v0 += 1
    return um


def ic_angle_cv2(roi: np.ndarray, x: int, y: int, um) -> float:
    m01 = m10 = 0
    r = roi.astype(np.int64)
    for u in range(-HALF_PATCH, HALF_PATCH + 1):
        m10 += u * int(r[y, x + u])
    for v in range(1, HALF_PATCH + 1):
        d = um[v]
        vs = 0
        for u in range(-d, d + 1):
            a, b = int(r[y + v, x + u]), int(r[y - v, x + u])
            vs += a - b
            m10 += u * (a + b)
        m01 += v * vs
    return float(cv2.fastAtan2(float(m01), float(m10)))


def descriptor_np(blur: np.ndarray, x: int, y: int, angle_deg: float, pattern: np.ndarray) -> np.ndarray:
    factor = f32(np.float64(math.pi) / np.float64(f32(180.0)))
    ang = f32(f32(angle_deg) * factor)
    a, b = f32(_libm.cosf(ang)), f32(_libm.sinf(ang))
    px = pattern[:, 0].astype(f32); py = pattern[:, 1].astype(f32)
    rr = np.rint((px * b).astype(f32) + (py * a).astype(f32)).astype(np.int64)
    cc = np.rint((px * a).astype(f32) - (py * b).astype(f32)).astype(np.int64)
    vals = blur[y + rr, x + cc].astype(np.int32)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8).reshape(32, 8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


def load_pattern() -> np.ndarray:
    from pathlib import Path
    txt = (Path(__file__).resolve().parent.parent / "oracle" / "orb_pattern_31.inc").read_text()
    vals = [int(t) for line in txt.splitlines() if not line.startswith("//") for t in line.split(",") if t.strip()]
    return np.array(vals, np.int32).reshape(512, 2)


def extract_cv2(img: np.ndarray, quadtree, nfeatures=2000, nlevels=8, sf=1.2, ini_th=12, min_th=7, quota=None):
    """ORBextractor::operator() (lapping {0,0}); `quadtree(kps, minX,maxX,minY,maxY,N)` is injected."""
    from oracle import KP_DTYPE
    sc, inv = scale_tables(nlevels, sf)
    _, rois = pyramid_cv2(img, nlevels, sf)
    um = umax_table()
    pattern = load_pattern()
    all_k, all_d, per_level = [], [], []
    for l in range(nlevels):
        roi = rois[l]
        h, w = roi.shape
        cand = cell_candidates_cv2(roi, ini_th, min_th)
        kin = np.zeros(len(cand), KP_DTYPE)
        kin["x"], kin["y"], kin["response"] = cand[:, 0], cand[:, 1], cand[:, 2]
        kin["size"], kin["angle"], kin["class_id"] = 7, -1, -1
        sel = quadtree(kin, 16, w - 16, 16, h - 16, int(quota[l]))
        sel["x"] += 16; sel["y"] += 16
        sel["octave"] = l
        sel["size"] = int(f32(31) * sc[l])
        for k in sel:
            k["angle"] = ic_angle_cv2(roi, int(k["x"]), int(k["y"]), um)
        per_level.append((cand, sel.copy()))
        if len(sel) == 0:
            continue
        blur = cv2.GaussianBlur(roi.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        desc = np.stack([descriptor_np(blur, int(k["x"]), int(k["y"]), float(k["angle"]), pattern) for k in sel])
        if l != 0:
            sel["x"] = (sel["x"] * sc[l]).astype(f32); sel["y"] = (sel["y"] * sc[l]).astype(f32)
        all_k.append(sel); all_d.append(desc)
    return np.concatenate(all_k), np.concatenate(all_d), per_level, rois


# ------------------------------------------------------------------------------------------------
# DepthModule with cv2 primitives
# ------------------------------------------------------------------------------------------------

def project_cv2(pts4xn: np.ndarray, P: np.ndarray, W: int, H: int, min_d=5.0, max_d=200.0) -> np.ndarray:
    raw = np.zeros((H, W), np.float32)
    Q = cv2.gemm(P.astype(np.float32), pts4xn.astype(np.float32), 1.0, None, 0.0)
    inv = cv2.divide(1.0, Q[2:3])
    u = cv2.multiply(Q[0:1], inv)[0]; v = cv2.multiply(Q[1:2], inv)[0]; d = Q[2]
    ok = (u > 0) & (v > 0) & (u < W) & (v < H) & (d > min_d) & (d < max_d)
    idx = np.nonzero(ok)[0]
    ui = u[idx].astype(np.int32); vi = v[idx].astype(np.int32)      # C truncation (values > 0)
    for k in range(len(idx)):                                       # sequential: later overwrites
        raw[vi[k], ui[k]] = d[idx[k]]
    return raw


def inverse_dilation_cv2(raw: np.ndarray, mask: np.ndarray, max_d=200.0, scale=1.0) -> np.ndarray:
    M = float(np.float32(max_d) * np.float32(scale))
    t = cv2.subtract(np.full(raw.shape, M, np.float32), raw)
    _, t = cv2.threshold(t, M - 1, 0, cv2.THRESH_TOZERO_INV)
    t = cv2.dilate(t, mask, anchor=(-1, -1), iterations=1)
    t = cv2.subtract(np.full(raw.shape, M, np.float32), t)
    _, t = cv2.threshold(t, M - 1, 0, cv2.THRESH_TOZERO_INV)
    return t
