"""Host-only paths of the drop-in `age_align` program (no alignment is run, so they work without a GPU): option
handling and the fatal errors of both dialects, with the reference's messages and exit codes
(age_align.cpp:99-147,205-221; Trash/age_align.cpp:97-111,160-302,365)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "asmvar_b200" / "age_align"


def run(args, stdin=None):
    return subprocess.run([str(EXE), *args], capture_output=True, text=True, input=stdin, timeout=120)


def test_version_original_dialect():
    p = run(["-version"])
    assert p.returncode == 0 and p.stdout.startswith("AGE ")


def test_batch_usage_exits_1():
    p = run(["-h", "-t", "x.fa", "-q", "y.fa"])
    assert p.returncode == 1 and "Usage :" in p.stderr and "--target" in p.stderr


def test_batch_requires_target_and_query():
    p = run(["-t", "only_target.fa", "var.txt"])
    assert p.returncode == 1 and "Query  Fa is required" in p.stderr


def test_batch_missing_fasta_is_fatal():
    p = run(["-t", "missing.fa", "-q", "missing.fa", "var.txt"])
    assert p.returncode == 1 and "Cannot open file : missing.fa" in p.stderr and p.stdout == ""


def test_batch_one_mode_only():
    p = run(["-i", "-d", "-t", "x.fa", "-q", "y.fa", "v.txt"])
    assert p.returncode == 1 and "More than one mode is specified" in p.stderr


def test_original_unknown_option_and_exit_code():
    p = run(["-bogus", "a.fa", "b.fa"])
    assert p.returncode == 1 and "Unknown option '-bogus'." in p.stderr and p.stdout == ""


def test_original_bad_coordinates_and_missing_second_file():
    p = run(["-indel", "-coor1=abc", "a.fa", "b.fa"])
    assert p.returncode == 1 and "Error in coordinate specification '-coor1=abc'." in p.stderr
    p = run(["-indel", "a.fa"])
    assert p.returncode == 1 and "No second file given." in p.stderr


def test_original_commands_from_stdin_are_parsed_per_line():
    p = run([], stdin="-indel -bogus a.fa b.fa\n-tdup -indel a.fa b.fa\n")
    assert p.returncode == 1
    assert "Unknown option '-bogus'." in p.stderr and "More than one mode is specified." in p.stderr


def test_age2text_host_errors(tmp_path):
    """the converter of the binary side channel needs no device: usage, and a file that is not a record stream"""
    exe = ROOT / "asmvar_b200" / "age2text"
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 1 and "Usage: age2text" in p.stderr
    bad = tmp_path / "x.agb"
    bad.write_bytes(b"not a record stream, but long enough to pass the size test")
    fa = tmp_path / "a.fa"
    fa.write_text(">a\nACGT\n")
    p = subprocess.run([str(exe), "-t", str(fa), "-q", str(fa), str(bad)], capture_output=True, text=True)
    assert p.returncode == 1 and "not an AGB1 file" in p.stderr
