#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
#
# Builds the *unmodified* reference AGE (bioinformatics-centre/AsmVar, src/AsmvarAGE)
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored).
# No reference source is copied into the repository: sources are staged in a scratch
# directory under oracle/_ref/stage, compiled, and the stage is deleted again.
#
# One build-only patch is applied to the staged copy: `traceBackMaxima` is declared
# `int` but has no return statement (AGEaligner.cpp:615-642); g++ >= 8 emits `ud2`
# there and the binary dies with SIGILL.  We append `return 0;` (the caller ignores
# the value, AGEaligner.cpp:563,597).
#
# Outputs:
#   oracle/_ref/age_align_ref        batch CLI (src/AsmvarAGE/src/age_align.cpp),  -O2, no OpenMP
#   oracle/_ref/age_align_ref_orig   original CLI (src/AsmvarAGE/src/Trash/age_align.cpp), -O2, no OpenMP
#   oracle/_ref/age_align_ref_omp    batch CLI with -O2 -fopenmp -DOMP (2 threads / alignment)
#   oracle/_ref/age_align_ref_shipped  batch CLI exactly as the reference's Makefile builds it (src/AsmvarAGE/src/Makefile:1-3:
#                                    no optimisation flag, -fopenmp -DOMP): the "as shipped" flavour of bench.py's cpu_baseline
#   oracle/_ref/age_detect_dump      oracle/detect_dump.cpp (ours) around the AsmvarDetect flavour of the aligner
#                                    (src/AsmvarDetect/AGEaligner.{h,cpp}): dumps AlignResult for golden vectors
#   oracle/_ref/age_recall_dump      oracle/recall_dump.cpp (ours) around src/AsmvarDetect/VarUnit.cpp:
#                                    VarUnit::ReAlignAndReCallVar text + re-called variants for golden vectors
set -euo pipefail
REF=${AGE_REFERENCE_ROOT:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
SRC="$REF/src/AsmvarAGE/src"
if [ ! -d "$SRC" ]; then
  echo "build_ref.sh: reference sources not found at $SRC (fine on the GPU box: prebuilt oracle/_ref is used)" >&2
  exit 0
fi
mkdir -p "$OUT"
STAGE="$OUT/stage"
rm -rf "$STAGE"; mkdir -p "$STAGE/batch" "$STAGE/orig"
patch_tbm() {  # insert 'return 0;' before the closing brace of traceBackMaxima
  python3 - "$1" <<'PY'
import re, sys
p = sys.argv[1]
s = open(p).read()
i = s.index('int AGEaligner::traceBackMaxima')
j = s.index('int AGEaligner::addBpoints')
body = s[i:j]
k = body.rindex('}')
body = body[:k] + '\treturn 0;\n' + body[k:]
open(p, 'w').write(s[:i] + body + s[j:])
PY
}
cp "$SRC"/{AGEaligner.cpp,AGEaligner.hh,Sequence.hh,Scorer.hh,age_align.cpp,Join.cpp,Join.h} "$STAGE/batch/"
cp "$SRC"/Backup/{Fa.h,Region.h,gzstream.cpp,gzstream.h} "$STAGE/batch/"
cp "$SRC"/Trash/{AGEaligner.cpp,AGEaligner.hh,Sequence.hh,Scorer.hh,age_align.cpp} "$STAGE/orig/"
patch_tbm "$STAGE/batch/AGEaligner.cpp"
patch_tbm "$STAGE/orig/AGEaligner.cpp"
FLAGS='-DAGE_VERSION="v0.8" -DAGE_TIME -w -O2'
( cd "$STAGE/batch" && g++ $FLAGS AGEaligner.cpp age_align.cpp Join.cpp gzstream.cpp -lz -o "$OUT/age_align_ref" )
( cd "$STAGE/batch" && g++ $FLAGS -fopenmp -DOMP AGEaligner.cpp age_align.cpp Join.cpp gzstream.cpp -lz -o "$OUT/age_align_ref_omp" )
( cd "$STAGE/batch" && g++ -DAGE_VERSION="v0.8" -DAGE_TIME -w -fopenmp -DOMP AGEaligner.cpp age_align.cpp Join.cpp gzstream.cpp -lz -o "$OUT/age_align_ref_shipped" )
( cd "$STAGE/orig"  && g++ $FLAGS AGEaligner.cpp age_align.cpp -o "$OUT/age_align_ref_orig" )
DET="$REF/src/AsmvarDetect"
if [ -f "$DET/AGEaligner.cpp" ]; then
  mkdir -p "$STAGE/detect"
  cp "$DET"/{AGEaligner.cpp,AGEaligner.h,Sequence.h,Scorer.h} "$STAGE/detect/"
  patch_tbm "$STAGE/detect/AGEaligner.cpp"
  ( cd "$STAGE/detect" && g++ -w -O2 -I. AGEaligner.cpp "$HERE/detect_dump.cpp" -o "$OUT/age_detect_dump" )
  # VarUnit::ReAlignAndReCallVar / class AgeAlignment (VarUnit.cpp) around the same aligner
  cp "$DET"/{VarUnit.cpp,VarUnit.h,AgeOption.h,Region.h,Region.cpp,Fa.h,utility.h,utility.cpp,gzstream.h,gzstream.cpp} "$STAGE/detect/" 2>/dev/null || true
  ( cd "$STAGE/detect" && g++ -w -O2 -I. AGEaligner.cpp VarUnit.cpp Region.cpp "$HERE/recall_dump.cpp" -o "$OUT/age_recall_dump" ) \
    || ( cd "$STAGE/detect" && g++ -w -O2 -I. AGEaligner.cpp VarUnit.cpp Region.cpp utility.cpp gzstream.cpp "$HERE/recall_dump.cpp" -lz -o "$OUT/age_recall_dump" )
fi
rm -rf "$STAGE"
echo "built: $(ls "$OUT")"
