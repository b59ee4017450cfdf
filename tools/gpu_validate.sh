#!/bin/bash
# The state that ships: full GPU test suite, smoke, default bench line, reference arm, launch lists
O=gpurun_out
python -m pytest tests -m gpu -q > $O/val_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/val_tests.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/val_bench.json 2> $O/val_bench.err; echo "bench rc=$?"; tail -c 300 $O/val_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/val_bench_ref.json 2> $O/val_bench_ref.err; echo "ref rc=$?"; cat $O/val_bench_ref.json | cut -c1-400
B="python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 2"
for w in c2 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/val_launches_$w.csv $B --workload $w --steps 3 > $O/val_ncu_$w.log 2>&1
done
cap() { # tag kernel-regex workload skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o /tmp/$1 $B --workload $3 --steps 1 > $O/val_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $O/val_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > $O/val_$1_details.txt 2>/dev/null
  echo "captured $1"
}
cap enum age_enum_kernel c2 2
cap blocks age_blocks_kernel c2 2
cap presweep age_presweep_kernel c2 4
AGE_CLI_PROF=1 python tools/cli_throughput.py -n 100640 > $O/val_cli.txt 2>&1; tail -3 $O/val_cli.txt | cut -c1-300
