#!/bin/bash
# round 2, run y: run_len with resident-only look-ahead (+ wide straight runs on the maxima planes): parity, per-pair profile, A/B
O=gpurun_out
for lib in b200; do
  timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > $O/r3d_tests.log 2>&1
  echo "tests rc=$? $(tail -1 $O/r3d_tests.log)"
done
for w in c2 c4 c5; do
  AGE_LIB=$PWD/asmvar_b200/libage_b200_prof.so python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 0 --steps 1 --workload $w > $O/r3d_prof_$w.json 2> $O/r3d_prof_$w.err
  grep "p2 prof" $O/r3d_prof_$w.err | head -6
done
B="python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 2"
run() { # tag lib workload steps
  AGE_LIB=$PWD/asmvar_b200/$2 $B --workload $3 --steps $4 > $O/r3d_$3_$1.json 2> $O/r3d_$3_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r3d_$3_$1.json").read().strip().splitlines()[-1])
    print("$1 $3", round(d["value"],1), round(d["ms_per_step"],2), round(d["roofline"]["kernel_ms"],2))
except Exception as e: print("$1 $3 failed", e)
PY
}
for w in c2 c4 c5; do run base libage_b200_base.so $w 4; run new libage_b200.so $w 4; done
