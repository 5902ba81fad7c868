"""Development aid: one local bundle adjustment on the device (for ncu launch lists)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import ba_data
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S
c = F.Context(S.KITTI_W, S.KITTI_H, 1000, max_batch=1)
p = ba_data.make_problem(42, n_kf=20, n_fixed=4, n_points=2500, outlier_frac=0.02)
for _ in range(2):
    t0 = time.perf_counter(); r = F.local_bundle_adjustment(c, *ba_data.args(p)); print("ms", 1e3 * (time.perf_counter() - t0), "iters", r[3])
c.close()
