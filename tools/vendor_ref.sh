#!/bin/bash
# Copies the reference's hot-path package (easykv/*.py, unmodified) into oracle/_ref/ so that reference code can
# travel to the GPU box with `gpurun` (the box has no /root/reference).  oracle/_ref/ is git-ignored — reference
# sources never enter the history — but NOT gpurun-ignored.  Test infrastructure only: tests/, smoke() and bench.py's
# reference legs are the only importers (oracle/ref_harness.py resolves /root/reference first, oracle/_ref second).
set -eu
SRC=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DST=$ROOT/oracle/_ref
[ -d "$SRC/easykv" ] || { echo "no reference at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST/easykv"
cp "$SRC"/easykv/*.py "$DST/easykv/"
( cd "$SRC" && { git rev-parse HEAD 2>/dev/null || echo a1d71cae3b562d9a709dda3741bd63e46a09ad31; } ) > "$DST/COMMIT"
( cd "$DST/easykv" && sha256sum *.py ) > "$DST/SHA256SUMS"
echo "vendored $(ls "$DST/easykv" | wc -l) files into $DST (commit $(cat "$DST/COMMIT"))"
