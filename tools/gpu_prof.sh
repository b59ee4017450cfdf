#!/bin/bash
# diagnostic build (tools/build_variant.sh prof -DAGE_P2_PROF -> libage_b200_prof.so): cycles of pass 2 per pair, K7 counters
O=gpurun_out; T=${1:-prof}
for w in c2 c4 c5; do
  AGE_LIB=$PWD/asmvar_b200/libage_b200_prof.so python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 0 --steps 1 --workload $w > $O/${T}_$w.json 2> $O/${T}_$w.err
  grep "p2 prof" $O/${T}_$w.err | head -9
done
