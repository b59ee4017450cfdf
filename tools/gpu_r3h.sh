#!/bin/bash
# round 2, run 3h: pipelined sweep with the boundary ring fed per step (flush every 8 / 16 / 32 blocks): parity, then speed
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > $O/r3h_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/r3h_tests.log)"
B="python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 2 --steps 4"
run() { # tag lib workload batch
  AGE_LIB=$PWD/asmvar_b200/$2 timeout 120 $B --workload $3 --batch $4 > $O/r3h_$3_$1.json 2> $O/r3h_$3_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r3h_$3_$1.json").read().strip().splitlines()[-1])
    print("$1 $3 batch $4:", round(d["value"],1), "realign/s", round(d["ms_per_step"],2), "ms, sweep", round(d["roofline"]["kernel_ms"],2), "frac", round(d["roofline"]["frac"],3))
except Exception as e: print("$1 $3 failed", e, open("$O/r3h_$3_$1.err").read()[-300:])
PY
}
for lib in pf:libage_b200_pf.so fl8:libage_b200.so fl16:libage_b200_fl16.so fl32:libage_b200_fl32.so; do
  run ${lib%%:*} ${lib##*:} c5 74; run ${lib%%:*} ${lib##*:} c4 296
done
run pf libage_b200_pf.so c2 7104; run fl8 libage_b200.so c2 7104
