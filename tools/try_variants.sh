#!/bin/bash
# Round-2 first step (needs a GPU: run under gpurun).  The kernels prepared at the end of round 1 are validated on the CPU
# through their host twins only; each is selected by an environment variable read when a context is created, so every variant
# runs in its OWN process (a faulting kernel cannot poison the other runs): parity tests first, then the frame-construction
# timings of bench.py.  Outputs under gpurun_out/variants/.
#   gpurun --timeout 1500 -- 'bash tools/try_variants.sh'
set -u
out=gpurun_out/variants; mkdir -p "$out"
run() {   # name, env assignments...
    local name=$1; shift
    echo "== $name"
    env "$@" timeout 600 python -m pytest tests/test_gpu_extractor.py tests/test_gpu_depth.py -x -q -m gpu > "$out/${name}_tests.log" 2>&1
    echo "   tests exit $? ($(tail -1 "$out/${name}_tests.log"))"
    env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-bow > "$out/${name}_bench.json" 2> "$out/${name}_bench.err"
    echo "   bench exit $?"
    python - "$out/${name}_bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["kernels"]
    print("   value %.0f fps, e2e %.0f fps; fast %.3f ms, quadtree %.3f ms, describe %.3f ms, dilate %.3f ms per step" % (
        d["value"], d["e2e"]["value"], k["fast"]["ms_per_step"], d["latency_bound_stages"]["quadtree"]["ms_per_step"], k["describe"]["ms_per_step"],
        k["depth_resolve_dilate"]["ms_per_step"]))
except Exception as e:
    print("   no bench line:", e)
PY
}
run baseline RGBL_NONE=0
run fast_strips RGBL_FAST_STRIPS=1
run describe_staged RGBL_DESCRIBE_STAGED=1
run qt_block_sort RGBL_QT_BLOCK_SORT=1
run dilate_v2 RGBL_DILATE_V2=1
run all RGBL_FAST_STRIPS=1 RGBL_DESCRIBE_STAGED=1 RGBL_QT_BLOCK_SORT=1 RGBL_DILATE_V2=1
# the LM variant is a compile-time switch: rebuild pose_kernels.o with it (one file, ~1 min), run the tracking tests, restore
(
    cd orb_slam3_rgbl_b200/csrc || exit 1
    F="-gencode arch=compute_100a,code=sm_100a --extended-lambda -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off -fmad=true"
    cp obj/pose_kernels.o /tmp/pose_kernels.o.keep && cp ../librgbl_b200.so /tmp/librgbl_b200.so.keep || exit 1
    nvcc $F -DPOSE_MIXED_SOLVE=1 -c -o obj/pose_kernels.o pose_kernels.cu && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../librgbl_b200.so obj/*.o
) > "$out/mixed_build.log" 2>&1 && {
    timeout 900 python -m pytest tests/test_gpu_tracking.py -x -q -m gpu > "$out/pose_mixed_tests.log" 2>&1
    echo "== pose_mixed_solve tests exit $? ($(tail -1 "$out/pose_mixed_tests.log"))"
    timeout 600 python bench.py --steps 10 --warmup 3 --no-bow > "$out/pose_mixed_bench.json" 2> "$out/pose_mixed_bench.err"
    tail -c 600 "$out/pose_mixed_bench.json"; echo
}
[ -f /tmp/pose_kernels.o.keep ] && cp /tmp/pose_kernels.o.keep orb_slam3_rgbl_b200/csrc/obj/pose_kernels.o
[ -f /tmp/librgbl_b200.so.keep ] && cp /tmp/librgbl_b200.so.keep orb_slam3_rgbl_b200/librgbl_b200.so
