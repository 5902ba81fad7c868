"""Development aid: one resident process + tracking chain (for ncu launch lists)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S
T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
seq = S.PlaneSequence(1000, T + 1)
imgs = [seq.image(t) for t in range(T)]; pcs = [seq.cloud(t) for t in range(T)]
c = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=T, max_points=pcs[0].shape[1])
b = F.RgblBatch(c, imgs, pcs, seq.P, F.make_depth_params(bf=S.KITTI_BF), pinned=False)
b.upload()
for _ in range(2):
    b.process_resident()
    poses, nm, ni = b.track(seq.pose(0), S.KITTI_FX, S.KITTI_FY, S.KITTI_CX, S.KITTI_CY, S.KITTI_BF)
print(nm, ni)
c.close()
