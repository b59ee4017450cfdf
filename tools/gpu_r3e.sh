#!/bin/bash
# round 2, run 3e: full GPU test suite, the default bench line, launch lists and ncu captures of the pass-2 kernels after the
# enumeration work
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r3e_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/r3e_tests.log)"
python bench.py > $O/r3e_bench.json 2> $O/r3e_bench.err; echo "bench rc=$?"; tail -c 600 $O/r3e_bench.err
B="python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 2"
for w in c2 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r3e_launches_$w.csv $B --workload $w --steps 3 > $O/r3e_ncu_$w.log 2>&1
done
cap() { # tag kernel-regex workload skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o /tmp/$1 $B --workload $3 --steps 1 > $O/r3e_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $O/r3e_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > $O/r3e_$1_details.txt 2>/dev/null
  echo "captured $1 rc=$?"
}
cap enum age_enum_kernel c2 2
cap blocks age_blocks_kernel c2 2
cap presweep age_presweep_kernel c2 4
cap enum_c4 age_enum_kernel c4 2
ls -la $O | grep r3e | head -40
