#!/bin/bash
# Round-2 (second session) measurement pass, same steps as tools/profile_r02.sh with the kernel set of the final tree (run under gpurun, 1 GPU): GPU parity tests, bench line, ncu launch list of the same command, and ONE
# `ncu --set full` capture per frame-construction / chain kernel.  Outputs under gpurun_out/r02/<tag>/ ; summarise here with
# tools/ncu_report_all.py and copy what should be judged into profiles/.
#   gpurun --timeout 1500 -- 'bash tools/profile_r02.sh <tag>'
set -u
tag=${1:-a}
out=gpurun_out/r02b/prof_$tag; mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/gpu.txt"; nproc >> "$out/gpu.txt"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1200 python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest exit $? ($(tail -1 "$out/pytest_gpu.log"))"
fi
timeout 600 python bench.py --steps 20 --warmup 3 > "$out/bench.json" 2> "$out/bench.err"; echo "bench exit $?"; tail -c 400 "$out/bench.json"; echo
PROF="python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu-baseline --no-bow --multi-sequences 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 500 --csv --log-file "$out/launches.csv" $PROF > "$out/launches_run.log" 2>&1
echo "launch list exit $?"
python tools/summarize_ncu_launches.py "$out/launches.csv" > "$out/launches_summary.csv" 2>/dev/null; head -20 "$out/launches_summary.csv"
# two passes, the .ncu-rep files stay on the box (tens of MB): only their raw pages (CSV, one row per launch) come back
K1='regex:(resize_level|fast_|cand_|blur_level|quadtree_kernel|sel_pack|describe|depth_project|depth_resolve|depth_gather|grid_build|level_tile)'
K2='regex:(search_last_collect|search_local_collect|resolve_kernel|pose_optimize|chain_prep|tlm_prepare)'
timeout 900 ncu --set full --clock-control none -k "$K1" -s 62 -c 31 -o /tmp/fc_kernels -f $PROF > "$out/ncu_full.log" 2>&1
echo "ncu full (frame construction) exit $?"
timeout 900 ncu --set full --clock-control none -k "$K2" -s 40 -c 16 -o /tmp/chain_kernels -f $PROF >> "$out/ncu_full.log" 2>&1
echo "ncu full (chain) exit $?"
ncu -i /tmp/fc_kernels.ncu-rep --page raw --csv > "$out/fc_kernels_raw.csv" 2>> "$out/ncu_full.log"
ncu -i /tmp/chain_kernels.ncu-rep --page raw --csv > "$out/chain_kernels_raw.csv" 2>> "$out/ncu_full.log"
python tools/ncu_report_all.py "$out/fc_kernels_raw.csv" --longest > "$out/ncu_full_all.txt" 2>> "$out/ncu_full.log"
python tools/ncu_report_all.py "$out/chain_kernels_raw.csv" --longest >> "$out/ncu_full_all.txt" 2>> "$out/ncu_full.log"
ls -la "$out"; du -sh gpurun_out
