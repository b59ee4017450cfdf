#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Regenerates the CLI goldens that oracle/make_golden.py does not write:
# tests/golden/cli/orig/{cmd,expected}{3..6} and tests/golden/cli/test2/{cmd,expected}{1..4} -- the outputs of the
# reference binaries (oracle/_ref, built by oracle/build_ref.sh) for -revcom1, -inv -both, -invl, -invr in both dialects.
# The command lines are the committed cmdN.txt files.  Run in the build container only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
CLI="$HERE/../tests/golden/cli"
for i in 3 4 5 6; do
  ( cd "$CLI/orig" && "$HERE/_ref/age_align_ref_orig" $(cat cmd$i.txt) 2>/dev/null | grep -v '^Alignment time is' > expected$i.age || true )
done
for i in 1 2 3 4; do
  ( cd "$CLI/test2" && "$HERE/_ref/age_align_ref" $(cat cmd$i.txt) 2>/dev/null | grep -v '^Alignment time is' > expected$i.age )
done
echo "regenerated $(ls "$CLI"/orig/expected[3-6].age "$CLI"/test2/expected[1-4].age | wc -l) files"
