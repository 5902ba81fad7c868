#!/bin/bash
# Development aid (run under gpurun): cluster shape sweep of pose_optimize_kernel (CTAs x threads), chain time + parity per shape.
set -e
cd orb_slam3_rgbl_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a --extended-lambda -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off -fmad=true"
cp obj/pose_kernels.o /tmp/pose_kernels.o.keep; cp ../librgbl_b200.so /tmp/librgbl_b200.so.keep
for cfg in "4 256" "8 128" "8 256" "16 64" "16 128"; do
  set -- $cfg
  nvcc $F -DPOSE_CTAS=$1 -DPOSE_THREADS=$2 -c -o obj/pose_kernels.o pose_kernels.cu 2>/dev/null && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../librgbl_b200.so obj/*.o
  cd ../..
  echo "== cluster $1 x $2"
  python -m pytest tests/test_gpu_tracking.py -x -q -m gpu 2>&1 | tail -1
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-side 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'chain', d['tracking_chain']['ms_per_step'], 'us/frame', d['tracking_chain']['us_per_frame'])"
  cd orb_slam3_rgbl_b200/csrc
done
cp /tmp/pose_kernels.o.keep obj/pose_kernels.o; cp /tmp/librgbl_b200.so.keep ../librgbl_b200.so
