#!/bin/bash
# round 2, run x: diagnostic build (-DAGE_P2_PROF): per-pair cycles of the enumerate / blocks kernels
O=gpurun_out
for w in c2 c4 c5; do
  AGE_LIB=$PWD/asmvar_b200/libage_b200_prof.so python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 0 --steps 1 --workload $w > $O/r2x_prof_$w.json 2> $O/r2x_prof_$w.err
  grep "p2 prof" $O/r2x_prof_$w.err | head -40
done
