"""Development aid: where does a step of rgbl_track_sequence go?  resident vs host inputs vs host inputs + frame outputs."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import bench
from orb_slam3_rgbl_b200 import frontend as F
cfg = bench.CONFIGS["B"]; T, M = cfg["T"], cfg["M"]
seq = bench.make_sequence(cfg, 1000)
ctx = F.Context(cfg["W"], cfg["H"], cfg["nfeat"], max_batch=T, max_points=seq.cloud(0).shape[1])
prm = F.make_depth_params(bf=cfg["cam"][4])
r = F.SequenceRunner(ctx, seq.P, prm, T, cfg["W"], cfg["H"], seq.cloud(0).shape[1], M, pinned=True)
for m in range(M):
    r.set_batch(m, [seq.image(m * T + f) for f in range(T)], [seq.cloud(m * T + f) for f in range(T)]); r.stage(m, m)
ch = lambda c, K=3: F.make_chain_params(seq.pose(0), *cfg["cam"], th_last=15.0, continue_sequence=c, local_map_frames=K, th_local=3.0)
r.run(ch(False), 3, resident_slots=M)
for name, kw in (("resident", dict(resident_slots=M)), ("host in", dict(resident_slots=0)), ("host in + frames out", dict(resident_slots=0, want_frames=True)),
                 ("resident K=0", dict(resident_slots=M))):
    K = 0 if "K=0" in name else 3
    if K == 0:
        r.run(ch(False, 0), 2, **kw)
    ctx.profile_enable(1); ctx.profile_reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r.run(ch(True, K), 10, **kw)
    torch.cuda.synchronize(); ms = 1e3 * (time.perf_counter() - t0) / 10
    p = ctx.profile_read(); ctx.profile_enable(0)
    print(f"{name:24s} {ms:7.2f} ms/step, chain {p['match']['ms'] / 10:7.2f} ms/step (device, begin->end events), launches/step {p['_total_launches'] / 10:.0f}")
ctx.close()
