"""rgbl_frame_rgbl_batch (extraction + LiDAR depth of a 2-frame batch) through the real C ABI of the EMULATED library
(tests/cuda_emu, `build.py full`), compared bit for bit with the oracle.  Run it once with the default kernels and once with every
prepared variant selected by its environment variable (this is what exercises the selection code of api.cu without a GPU):
    python tests/tools/emu_abi_variants.py
    RGBL_FAST_STRIPS=1 RGBL_DESCRIBE_STAGED=1 RGBL_QT_BLOCK_SORT=1 RGBL_DILATE_V2=1 python tests/tools/emu_abi_variants.py
"""
import sys, os
ROOT = __import__('pathlib').Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from pathlib import Path
from orb_slam3_rgbl_b200 import _lib
import importlib.util
spec = importlib.util.spec_from_file_location('cuda_emu_build', ROOT / 'tests' / 'cuda_emu' / 'build.py')
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
_lib.LIB_PATH = mod.build_full()
import numpy as np, oracle
from orb_slam3_rgbl_b200 import frontend as F, synthetic as S
W,H=420,260
img=S.make_image(5,W,H); pts=S.make_pointcloud(5,n_azimuth=500)
P=S.lidar_projection_matrix().copy(); P[0,:]*=W/S.KITTI_W; P[1,:]*=H/S.KITTI_H
ctx=F.Context(W,H,600,max_batch=2,max_points=pts.shape[1])
res=F.frame_rgbl_batch(ctx,[img,img[:, ::-1].copy()],[pts,pts],P,F.make_depth_params(bf=S.KITTI_BF))
ctx.close()
for (k,d,dep,ur),im in zip(res,[img,img[:, ::-1].copy()]):
    rk,rd,_=oracle.Extractor(600)(im)
    rdep,rur,_,_=oracle.depth_from_pcd(pts,P,W,H,S.structuring_element("diamond",5),S.KITTI_BF,rk,rk)
    ok = len(k)==len(rk) and all((k[f]==rk[f]).all() for f in k.dtype.names) and (d==rd).all() and (dep==rdep).all() and (ur==rur).all()
    print("frame ok" if ok else "MISMATCH", len(k), int((dep>0).sum()))
