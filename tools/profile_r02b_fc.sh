#!/bin/bash
# Short form of tools/profile_r02b.sh: the ncu launch list of the bench command and the `ncu --set full` pass over the frame-construction kernels only
# (the chain kernels did not change after the full pass).  gpurun --timeout 260 -- 'bash tools/profile_r02b_fc.sh <tag>'
set -u
tag=${1:-c}
out=gpurun_out/r02b/prof_$tag; mkdir -p "$out"
PROF="python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu-baseline --no-bow --multi-sequences 0"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 500 --csv --log-file "$out/launches.csv" $PROF > "$out/launches_run.log" 2>&1
python tools/summarize_ncu_launches.py "$out/launches.csv" > "$out/launches_summary.csv" 2>/dev/null; head -8 "$out/launches_summary.csv"
K1='regex:(fast_|cand_|quadtree_kernel|sel_pack|describe|depth_project|depth_resolve|depth_gather|grid_build|level_tile)'
timeout 200 ncu --set full --clock-control none -k "$K1" -s 62 -c 31 -o /tmp/fc_kernels -f $PROF > "$out/ncu_full.log" 2>&1
ncu -i /tmp/fc_kernels.ncu-rep --page raw --csv > "$out/fc_kernels_raw.csv" 2>> "$out/ncu_full.log"
ls -la "$out" | tail -5
