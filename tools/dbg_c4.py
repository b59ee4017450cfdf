import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asmvar_b200 import synth, engine, capi
c = synth.inversion_candidates(192, seed=1004)
with engine.Engine() as e:
    for i in [int(a) for a in sys.argv[1:]]:
        w, q = c[i]
        r = e.align([(w, q, capi.INVERSION_FLAG, 1, -1, -10, -1, -1)])
        st = e.stats()
        print("cand", i, "pass2_ms", round(st["kernel_ms"] - st["fill_ms"], 1), r[0]["bp"][:3])
