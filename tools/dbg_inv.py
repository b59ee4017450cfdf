import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ageutil as U
from asmvar_b200 import engine
cases = [c for c in U.load_cases() if c["flag"] in (2, 4, 8, 16)][:int(sys.argv[1]) if len(sys.argv) > 1 else 6]
with engine.Engine() as e:
    for c in cases:
        got = U.candidate_record(c, e.aligner_summary)
        print(c["id"], c["flag"], got.rstrip("\n") == c["expected"].rstrip("\n"))
