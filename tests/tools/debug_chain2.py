"""Development aid: prints the GPU chain (rgbl_resident_track_begin2) next to the oracle chain, field by field."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import tracking_data as TD
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T, nB = 5, 2
seq = S.PlaneSequence(31, T * nB + 1)
c = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=T, max_points=seq.cloud(0).shape[1])
prm = F.make_depth_params(bf=S.KITTI_BF)
got = []
for b in range(nB):
    B = F.RgblBatch(c, [seq.image(t) for t in range(b * T, (b + 1) * T)], [seq.cloud(t) for t in range(b * T, (b + 1) * T)], seq.P, prm, pinned=False)
    B.upload(); B.process_resident()
    B.track_begin2(F.make_chain_params(seq.pose(0), *TD.CAM, th_last=15.0, continue_sequence=b > 0, local_map_frames=K, th_local=3.0))
    got.append(B.track_end2())
c.close()
frames, sf = TD.extract_frames(seq, list(range(T * nB)))
state = None
for b in range(nB):
    rp, rnm, rni, rnl, rni1, state = TD.oracle_chain2(frames[b * T:(b + 1) * T], sf, seq.pose(0), K=K, state=state)
    g = got[b]
    print("batch", b)
    for name, ref in (("n_matches", rnm), ("n_inliers_first", rni1), ("n_local_matches", rnl), ("n_inliers", rni)):
        print("  %-16s gpu %s\n  %-16s ref %s" % (name, g[name], "", ref))
    print("  pose diff per frame", np.abs(g["poses"] - rp).max(1))
