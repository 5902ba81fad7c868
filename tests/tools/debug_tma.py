"""Debug aid: one extraction with the fused TMA level kernel, compared level by level with the two-pass kernels (run under
compute-sanitizer / CUDA_LAUNCH_BLOCKING=1 on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (S.KITTI_W, S.KITTI_H)
imgs = [S.make_image(60 + i, W, H, n_rects=150) for i in range(2)]
out = {}
for tma in ("0", "1"):
    os.environ["RGBL_LEVEL_TMA"] = tma
    ex = F.ORBextractor(1000, 1.2, 8, 12, 7, W, H, max_batch=2)
    ex.extract_batch(imgs)
    out[tma] = [[(ex.level_image(l, f), ex.blurred_level(l, f)) for l in range(8)] for f in range(2)]
    ex.ctx.close()
for f in range(2):
    for l in range(8):
        a, ab = out["1"][f][l]; b, bb = out["0"][f][l]
        dp, db = (a != b), (ab != bb)
        print(f"frame {f} level {l} {a.shape}: pyramid mismatches {dp.sum()} blur mismatches {db.sum()}", end="")
        if dp.any(): ys, xs = np.nonzero(dp); print(f"  pyr first at (y={ys[0]}, x={xs[0]}) rows {np.unique(ys)[:8]} cols {np.unique(xs)[:8]}", end="")
        if db.any(): ys, xs = np.nonzero(db); print(f"  blur first at (y={ys[0]}, x={xs[0]}) rows {np.unique(ys)[:8]} cols {np.unique(xs)[:8]}", end="")
        print()
