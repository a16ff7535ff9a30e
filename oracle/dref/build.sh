#!/bin/sh
# oracle/dref/build.sh -- TEST INFRASTRUCTURE: build the UNMODIFIED reference (dbox) and the harness that pins the oracle to it.
#
# Needs a D compiler (ldc2, dmd or gdc); the build image has none (see profiles/r02_d_toolchain_probe.txt), so this recipe
# is for any machine that has one.  The library sources are compiled where they lie under $REF/src (they import only
# druntime / Phobos; dub is NOT used: it would try to fetch the demo's GUI dependencies); nothing is copied and every output
# goes to oracle/_ref/ (git-ignored).  Result: oracle/_ref/reference_golden.json -- copy it to tests/golden/ and commit it;
# tests/test_oracle.py::test_oracle_matches_reference_golden then compares the oracle with it bit for bit.
set -eu
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${DBOX_REFERENCE:-/root/reference}"
OUT="$HERE/../_ref"
mkdir -p "$OUT"
SRCS="$(find "$REF/src" -name '*.d')"
if command -v ldc2 >/dev/null 2>&1; then
  ldc2 -O2 -release -I"$REF/src" -of="$OUT/dref_harness" -od="$OUT/obj" "$HERE/harness.d" $SRCS
elif command -v dmd >/dev/null 2>&1; then
  dmd -O -release -inline -I"$REF/src" -of"$OUT/dref_harness" -od"$OUT/obj" "$HERE/harness.d" $SRCS
elif command -v gdc >/dev/null 2>&1; then
  gdc -O2 -frelease -I"$REF/src" -o "$OUT/dref_harness" "$HERE/harness.d" $SRCS
else
  echo "no D compiler (ldc2 / dmd / gdc) on this machine: the oracle stays PARITY UNPINNED" >&2
  exit 3
fi
"$OUT/dref_harness" > "$OUT/reference_golden.json"
echo "wrote $OUT/reference_golden.json"
