"""Stage-by-stage GPU-vs-oracle diagnostic (development aid; run under gpurun)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np
import oracle
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S

W, H, NF = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (S.KITTI_W, S.KITTI_H, 2000)
img = S.make_image(0, W, H)
ref = oracle.Extractor(NF); rk, rd, _ = ref(img)
ex = F.ORBextractor(NF, 1.2, 8, 12, 7, W, H)
mono, k, d = ex(img)
print("n gpu", len(k), "n oracle", len(rk))
for l in range(8):
    a, b = ex.level_image(l), ref.level_image(l)
    nm = int((a != b).sum())
    msg = f"L{l} {a.shape} pyr mism {nm}"
    if nm:
        ys, xs = np.nonzero(a != b); msg += f" first (y={ys[0]},x={xs[0]}) gpu={a[ys[0],xs[0]]} ref={b[ys[0],xs[0]]} rows {np.unique(ys)[:5]} cols {np.unique(xs)[:5]}"
    ab, bb = ex.blurred_level(l), oracle.gaussian_blur7(a)
    nb = int((ab != bb).sum()); msg += f" | blur mism {nb}"
    if nb:
        ys, xs = np.nonzero(ab != bb); msg += f" first (y={ys[0]},x={xs[0]}) gpu={ab[ys[0],xs[0]]} ref={bb[ys[0],xs[0]]}"
    ca, cb = ex.level_candidates(l), ref.level_candidates(l)
    sa = {tuple(r) for r in ca.tolist()}; sb = {tuple(r) for r in cb.tolist()}
    msg += f" | cand {len(ca)} vs {len(cb)} gpu-only {len(sa - sb)} ref-only {len(sb - sa)}"
    if len(ca) == len(cb) and sa == sb:
        msg += f" order-equal {bool((ca == cb).all())}"
    else:
        msg += f" e.g. gpu-only {sorted(sa - sb)[:3]} ref-only {sorted(sb - sa)[:3]}"
    print(msg)
n = min(len(k), len(rk))
for f in k.dtype.names:
    print(f, int((k[f][:n] != rk[f][:n]).sum()))
print("desc rows differing", int((d[:n] != rd[:n]).any(axis=1).sum()))
ca, cb = ex.level_candidates(0), ref.level_candidates(0)
print("gpu first 16:", ca[:16].tolist())
print("ref first 16:", cb[:16].tolist())
# per-cell view of the first cell row: window x in [0, wCell+6)
lv = ref.level_image(0)
import collections
def cell_hist(c, wc=36, hc=35):
    return collections.Counter(((r[1] - 3) // hc, (r[0] - 3) // wc) for r in c.tolist())
ha, hb = cell_hist(ca), cell_hist(cb)
print("cells (row,col)->count gpu:", sorted(ha.items())[:12])
print("cells (row,col)->count ref:", sorted(hb.items())[:12])
