#!/bin/bash
# Round-2 (second session) check, run under gpurun on 1 GPU: GPU parity tests, the bench line, in-stream chain timing and the
# clock64 phase breakdowns of resolve_kernel / pose_optimize_kernel (tools/dbg/librgbl_b200_dbg.so = the same sources built with
# -DRESOLVE_DEBUG -DPOSE_TIMING, built in the container).  Outputs under gpurun_out/r02b/<tag>/.
#   gpurun --timeout 900 -- 'bash tools/r02b_check.sh <tag>'
set -u
tag=${1:-a}
out=gpurun_out/r02b/$tag; mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/gpu.txt"; nproc >> "$out/gpu.txt"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest exit $? ($(tail -1 "$out/pytest_gpu.log"))"
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 600 python bench.py --steps 20 --warmup 3 > "$out/bench.json" 2> "$out/bench.err"; echo "bench exit $?"
  python - "$out/bench.json" <<'EOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3))
    for k, v in d.get("kernels", {}).items():
        print("  %-24s %.4f ms" % (k, v["ms_per_step"]))
    print("  frame construction serial", d.get("frame_construction", {}).get("ms_per_step_serial"))
    print("  chain", json.dumps(d.get("tracking_chain", {}))[:600])
except Exception as e:
    print("bench parse failed", e)
EOF
fi
if [ "${AB_PDL:-0}" = "1" ]; then
  RGBL_CHAIN_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bow --multi-sequences 0 > "$out/bench_nopdl.json" 2>/dev/null
  python -c "import json,sys; d=json.loads(open('$out/bench_nopdl.json').read().strip().splitlines()[-1]); print('RGBL_CHAIN_PDL=0: value', round(d['value'],1), 'chain us/frame', round(d['tracking_chain']['us_per_frame'],1))"
fi
RGBL_CHAIN_TIMING=1 timeout 300 python tools/run_chain_tlm.py 8 > "$out/chain_timing.log" 2>&1; echo "chain timing exit $?"; grep "chain timing" "$out/chain_timing.log" | tail -6
if [ -f tools/dbg/librgbl_b200_dbg.so ]; then
  cp orb_slam3_rgbl_b200/librgbl_b200.so /tmp/librgbl_b200.so.keep
  cp tools/dbg/librgbl_b200_dbg.so orb_slam3_rgbl_b200/librgbl_b200.so
  RGBL_CHAIN_GRAPH=0 timeout 300 python tools/run_chain_tlm.py 8 > "$out/phase_debug.log" 2>&1; echo "phase debug exit $?"
  cp /tmp/librgbl_b200.so.keep orb_slam3_rgbl_b200/librgbl_b200.so
  grep "^resolve" "$out/phase_debug.log" | tail -6
  grep "pose n=" "$out/phase_debug.log" | tail -4
fi
ls -la "$out"
