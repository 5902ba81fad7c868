"""Pins oracle/png_oracle.cpp (PNG reader + cvtColor to gray = cv::imread(IMREAD_UNCHANGED) at Examples/RGB-L/rgbl_kitti.cc:87 followed by
src/Tracking.cc:1567-1580): against python-cv2's libpng reader and cvtColor live when cv2 is importable, and against the committed
fixture tests/golden/png_golden.npz (made by tests/golden/make_png_golden.py with cv2 4.13.0) always."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from orb_slam3_rgbl_b200 import synthetic as S

GOLD = Path(__file__).resolve().parent / "golden" / "png_golden.npz"


def cases():
    g = S.make_image(3, 161, 101, n_rects=20)
    rng = np.random.default_rng(5)
    out = []
    for name, img in (("gray", g), ("rgb", S.colorize(g, 1)), ("rgba", S.colorize(g, 2, alpha=True))):
        out.append((name + "_cycle", S.encode_png(img, None, idat_chunk=5000)))
        out.append((name + "_random", S.encode_png(img, rng.integers(0, 5, img.shape[0]), idat_chunk=1 << 16)))
        for f in range(5):
            out.append((f"{name}_f{f}", S.encode_png(img[:37, :53], np.full(37, f), idat_chunk=700)))
    return out


def cv2_reference(png, camera_rgb):
    import cv2
    m = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_UNCHANGED)
    if m.ndim == 2:
        return m
    code = {(3, True): cv2.COLOR_RGB2GRAY, (3, False): cv2.COLOR_BGR2GRAY, (4, True): cv2.COLOR_RGBA2GRAY, (4, False): cv2.COLOR_BGRA2GRAY}
    return cv2.cvtColor(m, code[(m.shape[2], camera_rgb)])


def test_png_oracle_equals_cv2_live():
    pytest.importorskip("cv2")
    for name, png in cases():
        for rgb in (True, False):
            gray, _ = oracle.png_decode_gray(png, rgb)
            ref = cv2_reference(png, rgb)
            assert gray.shape == ref.shape and (gray == ref).all(), (name, rgb, int((gray != ref).sum()))


def test_png_oracle_equals_committed_cv2_fixture():
    gold = np.load(GOLD)
    for name, png in cases():
        assert bytes(gold[name + "_png"]) == png, "the synthetic PNG writer changed: regenerate tests/golden/png_golden.npz"
        for rgb in (True, False):
            gray, _ = oracle.png_decode_gray(png, rgb)
            assert (gray == gold[f"{name}_gray{int(rgb)}"]).all(), (name, rgb)


def test_png_oracle_rejects_what_the_path_does_not_take():
    g = S.make_image(3, 64, 48, n_rects=5)
    png = bytearray(S.encode_png(g))
    with pytest.raises(ValueError):
        oracle.png_decode_gray(bytes(png[:40]))                 # truncated
    bad = bytearray(png); bad[30] ^= 0xff                        # CRC of IHDR no longer matches
    with pytest.raises(ValueError):
        oracle.png_decode_gray(bytes(bad))
    with pytest.raises(ValueError):
        oracle.png_decode_gray(b"not a png at all, just bytes" * 4)
