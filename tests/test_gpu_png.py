"""GPU parity of the image input step: PNG file bytes -> gray level-0 images (rgbl_decode_png_gray / rgbl_resident_upload_kitti_png =
cv::imread(IMREAD_UNCHANGED) at Examples/RGB-L/rgbl_kitti.cc:87 + cvtColor at src/Tracking.cc:1567-1580) against the oracle
(oracle/png_oracle.cpp, pinned against cv2's libpng reader and cvtColor in tests/test_oracle_png.py).  Bar: bit-exact."""
import numpy as np
import pytest

import oracle
from orb_slam3_rgbl_b200 import frontend as F
from orb_slam3_rgbl_b200 import synthetic as S
from orb_slam3_rgbl_b200._lib import RgblError

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,kind", [(S.KITTI_W, S.KITTI_H, "rgb"), (S.KITTI_W, S.KITTI_H, "gray"), (641, 481, "rgba"), (1920, 1080, "rgb"), (333, 1100, "rgb")])
def test_png_decode_equals_oracle(W, H, kind):
    """every filter type (cycling and random per row), several IDAT chunk sizes, one band (H <= 512) and several bands (H = 1080, 1100)"""
    rng = np.random.default_rng(7)
    grays = [S.make_image(80 + i, W, H, n_rects=60) for i in range(3)]
    imgs = [g if kind == "gray" else S.colorize(g, i, alpha=(kind == "rgba")) for i, g in enumerate(grays)]
    pngs = [S.encode_png(imgs[0], None, idat_chunk=8192), S.encode_png(imgs[1], rng.integers(0, 5, H)), S.encode_png(imgs[2], np.full(H, 4), idat_chunk=100000)]
    ctx = F.Context(W, H, 1000, max_batch=3)
    try:
        for rgb in (True, False):
            got = F.decode_png_gray(ctx, pngs, rgb)
            for png, g in zip(pngs, got):
                ref, _ = oracle.png_decode_gray(png, rgb)
                assert (g == ref).all(), (kind, rgb, int((g != ref).sum()))
    finally:
        ctx.close()


def test_png_upload_gives_the_same_frames_as_gray_upload():
    """rgbl_resident_upload_kitti_png + process == rgbl_resident_upload_kitti of the oracle-decoded gray images + process"""
    T = 3
    seq = S.PlaneSequence(2000, T + 1)
    colour = [S.colorize(seq.image(t), t) for t in range(T)]
    pngs = [S.encode_png(c) for c in colour]
    grays = [oracle.png_decode_gray(p, True)[0] for p in pngs]
    clouds = [seq.cloud(t) for t in range(T)]
    xyzr = [np.ascontiguousarray(np.vstack([c[:3], np.zeros((1, c.shape[1]), np.float32)]).T) for c in clouds]
    ctx = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=T, max_points=max(c.shape[1] for c in clouds))
    try:
        outs = []
        for use_png in (True, False):
            b = F.RgblBatch(ctx, grays, clouds, seq.P, F.make_depth_params(bf=S.KITTI_BF), pinned=False)
            if use_png:
                b.upload_kitti_png(pngs, xyzr, True)
            else:
                b.upload_kitti(xyzr)
            b.process_resident()
            outs.append(b.download())
        for (k0, d0, z0, u0), (k1, d1, z1, u1) in zip(*outs):
            assert len(k0) == len(k1) and len(k0) > 500
            assert all((k0[f] == k1[f]).all() for f in k0.dtype.names) and (d0 == d1).all() and (z0 == z1).all() and (u0 == u1).all()
    finally:
        ctx.close()


def test_png_errors_are_reported():
    g = S.make_image(1, 320, 240, n_rects=10)
    ctx = F.Context(320, 240, 500, max_batch=1)
    try:
        png = S.encode_png(g)
        with pytest.raises(RgblError):
            F.decode_png_gray(ctx, [png[:100]])                      # truncated
        bad = bytearray(png); bad[len(bad) // 2] ^= 0x55
        with pytest.raises(RgblError):
            F.decode_png_gray(ctx, [bytes(bad)])                     # CRC / zlib corruption
        with pytest.raises(RgblError):
            F.decode_png_gray(ctx, [S.encode_png(g[:100, :100])])    # size mismatch
        assert (F.decode_png_gray(ctx, [png])[0] == g).all()         # the context still works afterwards
    finally:
        ctx.close()
