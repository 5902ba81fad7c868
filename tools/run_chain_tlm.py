"""Development aid: one KITTI-size batch through frame construction + the resident chain WITH TrackLocalMap (local_map_frames = 3),
twice (the second chain continues the sequence, so every frame has a full local map).  Used with RGBL_CHAIN_TIMING=1 (in-stream CUDA
event times of the middle frame, stderr) and with the -DRESOLVE_DEBUG / -DPOSE_TIMING builds (clock64 phase breakdowns, stdout)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seq = S.PlaneSequence(1000, 2 * T + 1)
c = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=T, max_points=seq.cloud(0).shape[1])
print("keypoint capacity per frame:", c.cap)
out = None
for rep in range(2):
    imgs = [seq.image(rep * T + t) for t in range(T)]; pcs = [seq.cloud(rep * T + t) for t in range(T)]
    if rep == 0:
        b = F.RgblBatch(c, imgs, pcs, seq.P, F.make_depth_params(bf=S.KITTI_BF), pinned=False)
    else:
        b.set_inputs(imgs, pcs)
    b.upload()
    n = b.process_resident()
    prm = F.make_chain_params(seq.pose(0), S.KITTI_FX, S.KITTI_FY, S.KITTI_CX, S.KITTI_CY, S.KITTI_BF, continue_sequence=(rep > 0), local_map_frames=3)
    b.track_begin2(prm)
    out = b.track_end2()
print("keypoints", n.tolist())
print("n_matches", out["n_matches"].tolist(), "n_local", out["n_local_matches"].tolist(), "n_inliers", out["n_inliers"].tolist())
c.close()
