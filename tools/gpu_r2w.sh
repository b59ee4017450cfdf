#!/bin/bash
# round 2, run w: bisect the pass-2 package (build variants v1..v8 of age_pass2.cuh's AGE_V_* switches) for parity and speed
set -u
O=gpurun_out
mkdir -p $O
for lib in base v1 v2 v7; do
  AGE_LIB=$PWD/asmvar_b200/libage_b200_$lib.so timeout 300 python -m pytest tests/test_gpu_parity.py -k other_kernel_paths -q > $O/r2w_knobs_$lib.log 2>&1
  echo "knobs $lib rc=$? $(tail -1 $O/r2w_knobs_$lib.log)"
done
AGE_LIB=$PWD/asmvar_b200/libage_b200_v7.so timeout 420 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -k "other_kernel_paths and knob4" -x -q > $O/r2w_memcheck_v7.log 2>&1
echo "memcheck rc=$?"; grep -m1 -A12 "Invalid\|Error" $O/r2w_memcheck_v7.log | head -30
B="python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 2"
run() { # tag workload steps
  AGE_LIB=$PWD/asmvar_b200/libage_b200_$1.so $B --workload $2 --steps $3 > $O/r2w_$2_$1.json 2> $O/r2w_$2_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2w_$2_$1.json").read().strip().splitlines()[-1])
    print("$1 $2", round(d["value"],1), round(d["ms_per_step"],2), round(d["roofline"]["kernel_ms"],2))
except Exception as e: print("$1 $2 failed", e)
PY
}
for lib in base v1 v2 v3 v4 v5 v6 v7 v8; do run $lib c2 4; run $lib c4 4; run $lib c5 3; done
