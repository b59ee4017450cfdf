#!/bin/bash
# round 2, run v: pass-2 latency package (page table as residency map, wide straight runs, compact kernels) -- parity + A/B
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2v_tests.log
B="python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 3"
run() { # tag lib workload steps
  AGE_LIB=$PWD/asmvar_b200/$2 $B --workload $3 --steps $4 > $O/r2v_$3_$1.json 2> $O/r2v_$3_$1.err
  python - <<PY
import json
d=json.loads(open("$O/r2v_$3_$1.json").read().strip().splitlines()[-1])
print("$1 $3", round(d["value"],1), round(d["ms_per_step"],2), round(d["roofline"]["kernel_ms"],2), d["roofline"]["frac"])
PY
}
run base libage_b200_base.so c2 6
run new  libage_b200.so c2 6
run occ5 libage_b200_p2occ5.so c2 6
run new  libage_b200.so c4 4
run occ5 libage_b200_p2occ5.so c4 4
run new  libage_b200.so c5 4
run occ5 libage_b200_p2occ5.so c5 4
for w in c2 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2v_launches_$w.csv $B --workload $w --steps 2 > $O/r2v_ncu_$w.log 2>&1
done
ls -la $O | grep r2v
