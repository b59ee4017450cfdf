"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/age_b200.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from asmvar_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    header = (ROOT / "include" / "age_b200.h").read_text()
    declared = set(re.findall(r"^[a-z_ ]*[ \*](age_[a-z_0-9]+)\(", header, re.M))
    assert declared == set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    # compile-time layout of the header as seen by a C compiler vs the ctypes mirror
    import subprocess, tempfile
    src = '#include "age_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",sizeof(age_job),sizeof(age_block),sizeof(age_result),sizeof(age_seqview),sizeof(age_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        p = Path(d) / "s.c"
        p.write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(p), "-o", str(Path(d) / "s")], check=True)
        out = subprocess.run([str(Path(d) / "s")], check=True, capture_output=True, text=True).stdout.split()
    want = [C.sizeof(capi.AgeJob), C.sizeof(capi.AgeBlock), C.sizeof(capi.AgeResult), C.sizeof(capi.AgeSeqView), C.sizeof(capi.AgeStats)]
    assert list(map(int, out)) == want


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = capi.load()
    ctx = lib.age_create(None, 0)
    assert not ctx
    assert b"no usable CUDA device" in lib.age_last_error(None)
    from asmvar_b200 import engine
    with pytest.raises(engine.AgeError):
        engine.Engine()


def test_host_helpers_match_reference_semantics():
    lib = capi.load()
    buf = C.create_string_buffer(8)
    lib.age_revcom(b"ACGTUNac", 8, buf)
    assert buf.raw == b"gtNAACGT"                       # Sequence.hh:138-150: U -> A, N unchanged, case kept
    assert lib.age_subst_score(b"A", b"a", 3, -2) == 3   # Sequence.hh:182-188: case-insensitive
    assert lib.age_subst_score(b"T", b"U", 3, -2) == -2  # U only matches U
    assert lib.age_subst_score(b"N", b"A", 3, -2) == 0   # Scorer.hh:28-31
