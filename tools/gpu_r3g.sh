#!/bin/bash
# round 2, run 3g: warps per pair and cluster size of the pipelined sweep on C5 / C4 (knobs AGE_LONG_NW / AGE_LONG_CS, any value now)
O=gpurun_out
B="python bench.py --no-cpu --no-cli --no-extra --no-inproc --warmup 2 --steps 4"
run() { # tag workload batch nw cs
  AGE_LONG_NW=$4 AGE_LONG_CS=$5 $B --workload $2 --batch $3 > $O/r3g_$1.json 2> $O/r3g_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r3g_$1.json").read().strip().splitlines()[-1])
    print("$1: $2 batch $3 nw $4 cs $5:", round(d["value"],1), "realign/s", round(d["ms_per_step"],2), "ms, sweep", round(d["roofline"]["kernel_ms"],2), "frac", round(d["roofline"]["frac"],3))
except Exception as e: print("$1 failed", e, open("$O/r3g_$1.err").read()[-300:])
PY
}
run a c5 72 0 0
run b c5 49 14 3
run c c5 98 14 3
run d c5 37 10 4
run e c5 74 10 4
run f c5 49 16 3
run g c5 72 14 2
run h c5 74 16 2
run i c4 296 0 0
run j c4 296 16 1
run k c4 148 16 1
run l c4 444 5 1
