#!/bin/sh
# Stage the UNMODIFIED reference package for timing on the GPU box's host cores (bench.py: cpu_baseline kind
# "reference").  TEST INFRASTRUCTURE ONLY: the reference is pure Python / PyTorch (no native code to compile), so
# "building" it means copying its seven library files from the read-only tree into oracle/_ref/, which is
# git-ignored (never committed) but travels to the GPU box with gpurun, where /root/reference does not exist.
# Nothing under smcdet_b200/ imports it; only bench.py's CPU-baseline leg does.
set -e
SRC="${SMCDET_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
if [ ! -d "$SRC/smcdet" ]; then
    echo "make_ref.sh: $SRC/smcdet not found (expected in the build container only); nothing staged" >&2
    exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref/smcdet"
cp "$SRC"/smcdet/*.py "$HERE/_ref/smcdet/"
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$HERE/_ref/REVISION"
echo "staged $(ls "$HERE/_ref/smcdet" | wc -l) files of the reference into $HERE/_ref/smcdet"
