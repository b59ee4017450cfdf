#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/align_result_cases.json: the AlignResult
(src/AsmvarDetect/AGEaligner.h:91-150) that the unmodified AsmvarDetect aligner (oracle/_ref/age_detect_dump, built by
oracle/build_ref.sh from /root/reference) produces for the inputs of tests/golden/aligner_cases.json -- for -both cases
also for the reverse-complemented contig (the second aligner of VarUnit.cpp:283-299).  Run in the build container only."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
import ageutil as U  # noqa: E402


def main():
    cases = U.load_cases()
    lines, meta = [], []
    for c in cases:
        variants = [("", c["s2"])] + ([("/rc", U.revcom_view(c["s2"]))] if c["both"] else [])
        for suffix, s2 in variants:
            s1 = c["s1"]
            if not s1["seq"] or not s2["seq"]:
                continue
            cid = c["id"] + suffix
            f = [cid, s1["name"], s1["start"], int(s1["reverse"]), s1["seq"], s2["name"], s2["start"], int(s2["reverse"]), s2["seq"],
                 c["flag"], c["match"], c["mismatch"], c["go"], c["ge"]]
            assert not any("\t" in str(x) or "\n" in str(x) for x in f)
            lines.append("\t".join(str(x) for x in f))
            meta.append({"id": cid, "case": c["id"], "rc": bool(suffix)})
    p = subprocess.run([str(ROOT / "oracle" / "_ref" / "age_detect_dump")], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True)
    out = [json.loads(l) for l in p.stdout.splitlines() if l.strip()]
    assert len(out) == len(meta), (len(out), len(meta))
    for o, m in zip(out, meta):
        assert o["id"] == m["id"]
        o.update(case=m["case"], rc=m["rc"])
    dst = ROOT / "tests" / "golden" / "align_result_cases.json"
    dst.write_text(json.dumps({"generator": "oracle/make_golden_detect.py (oracle/_ref/age_detect_dump = src/AsmvarDetect/AGEaligner.cpp, unmodified)",
                               "results": out}, separators=(",", ":")))
    print(f"{len(out)} AlignResult records -> {dst} ({dst.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
