"""Development aid: host-side phase times of the pipelined end-to-end loop (run_e2e | track_end | track_begin)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S
T = 32
seq = S.PlaneSequence(1000, T + 1)
imgs = [seq.image(t) for t in range(T)]; pcs = [seq.cloud(t) for t in range(T)]
c = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=T, max_points=pcs[0].shape[1])
b = F.RgblBatch(c, imgs, pcs, seq.P, F.make_depth_params(bf=S.KITTI_BF), pinned=True)
CAM = (S.KITTI_FX, S.KITTI_FY, S.KITTI_CX, S.KITTI_CY, S.KITTI_BF)
pose0 = seq.pose(0)
acc = np.zeros(3); pend = False; K = 12
c.profile_enable(True); c.profile_reset()
for it in range(K + 3):
    t0 = time.perf_counter(); b.run_e2e(); t1 = time.perf_counter()
    if pend: b.track_end()
    t2 = time.perf_counter(); b.track_begin(pose0, *CAM); pend = True; t3 = time.perf_counter()
    if it >= 3: acc += [t1 - t0, t2 - t1, t3 - t2]
b.track_end()
print("per step ms: run_e2e %.2f  track_end %.2f  track_begin %.2f  sum %.2f" % (*(1e3 * acc / K), 1e3 * acc.sum() / K))
pr = c.profile_read(); print("  chain (CUDA events on its stream): %.2f ms per call" % (pr["match"]["ms"] / pr["match"]["calls"]))
# same loop, resident
b.upload(); acc[:] = 0; pend = False; c.profile_reset()
for it in range(K + 3):
    t0 = time.perf_counter(); b.process_resident(); t1 = time.perf_counter()
    if pend: b.track_end()
    t2 = time.perf_counter(); b.track_begin(pose0, *CAM); pend = True; t3 = time.perf_counter()
    if it >= 3: acc += [t1 - t0, t2 - t1, t3 - t2]
b.track_end()
print("per step ms: process   %.2f  track_end %.2f  track_begin %.2f  sum %.2f" % (*(1e3 * acc / K), 1e3 * acc.sum() / K))
pr = c.profile_read(); print("  chain (CUDA events on its stream): %.2f ms per call" % (pr["match"]["ms"] / pr["match"]["calls"]))
# e2e alone (no chain)
t0 = time.perf_counter()
for it in range(K): b.run_e2e()
print("run_e2e alone %.2f ms" % (1e3 * (time.perf_counter() - t0) / K))
c.close()
