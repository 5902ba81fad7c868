#!/bin/bash
# ONE `ncu --set full` capture of the kernels matching a regex (source-level counters included), from a short bench run; the summary
# and the hottest source lines come back as text, the .ncu-rep stays on the box.
#   gpurun --timeout 600 -- 'bash tools/ncu_one.sh <tag> <kernel-regex> [launch-skip] [launch-count] [extra bench args]'
set -u
tag=$1; regex=$2; skip=${3:-0}; count=${4:-8}; shift 4 2>/dev/null || shift $#
out=gpurun_out/r02/$tag; mkdir -p "$out"
PROF="python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu-baseline --no-bow --multi-sequences 0 $*"
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:$regex" -s "$skip" -c "$count" -o /tmp/one -f $PROF > "$out/ncu.log" 2>&1
echo "ncu exit $?"
ncu -i /tmp/one.ncu-rep --page raw --csv > "$out/raw.csv" 2>> "$out/ncu.log"
python tools/ncu_report_all.py "$out/raw.csv" --longest > "$out/summary.txt" 2>> "$out/ncu.log"
ncu -i /tmp/one.ncu-rep --page source --csv --print-source sass > "$out/source_sass.csv" 2>> "$out/ncu.log"
ncu -i /tmp/one.ncu-rep --page source --csv --print-source cuda > "$out/source_cuda.csv" 2>> "$out/ncu.log" || true
cat "$out/summary.txt"; ls -la "$out"
