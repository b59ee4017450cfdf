#!/bin/bash
# round 2, run 3f: blocks kernel with staged blocks (one traversal), compact re-sweep with its loads ahead (build switch): parity + per-kernel times
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_recall.py -x -q > $O/r3f_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/r3f_tests.log)"
B="python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 2"
for lib in b200 b200_pf; do for w in c2 c4 c5; do
  AGE_LIB=$PWD/asmvar_b200/libage_$lib.so ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r3f_launches_${w}_$lib.csv $B --workload $w --steps 3 > $O/r3f_ncu_${w}_$lib.log 2>&1
done; done
python - <<'PY'
import csv,collections
for lib in ('b200','b200_pf'):
  for w in ('c2','c4','c5'):
    rows=[r for r in csv.reader(open(f'gpurun_out/r3f_launches_{w}_{lib}.csv')) if len(r)>10 and r[0].isdigit()]
    dd=collections.defaultdict(list)
    for r in rows: dd[r[4].split('(')[0].replace('void ','').replace('age_','')].append(int(r[-1])/1e6)
    print(lib, w, {n:[round(x,2) for x in sorted(set(round(y,2) for y in v))[:6]] for n,v in dd.items() if 'sweep_k' not in n and 'pack' not in n and 'long' not in n})
PY
