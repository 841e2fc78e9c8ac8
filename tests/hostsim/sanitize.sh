#!/usr/bin/env bash
# Memory-safety and race checks of the CUDA kernels' SOURCE on the CPU (compute-sanitizer is closed on the
# GPU pool): builds the hostsim emulation library with AddressSanitizer / ThreadSanitizer and runs the
# hostsim parity tests under it.  Host arrays are numpy buffers of the exact size and the emulated shared
# memory is an exactly sized heap block, so any out-of-bounds global or shared access of a kernel is reported;
# __syncthreads / __syncwarp / shuffles are std::barrier operations, so missing synchronisation shows up
# as a data race.  Known false positives: glibc lgammaf() writing `signgam` (CUDA's does not), and the
# uninstrumented OpenMP runtime of the oracle.
#   usage: tests/hostsim/sanitize.sh asan|tsan [pytest -k expression]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(cd "$HERE/../.." && pwd)"
KIND="${1:-asan}"; KEXPR="${2:-}"
GXX=/usr/bin/g++; [ -x "$GXX" ] || GXX=g++
if [ "$KIND" = asan ]; then SAN=address; RT=$($GXX -print-file-name=libasan.so); export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1
else SAN=thread; RT=$($GXX -print-file-name=libtsan.so); export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0 log_path=/tmp/smcdet_tsan"; fi
OUT=/tmp/libsmcdet_hostsim_$KIND.so
$GXX -std=c++20 -O1 -g -fsanitize=$SAN -fno-omit-frame-pointer -fPIC -shared -pthread -I"$HERE" -x c++ "$HERE/hostsim_lib.cpp" -o "$OUT"
cd "$ROOT"
SMCDET_HOSTSIM_LIB="$OUT" LD_PRELOAD="$RT" python -m pytest tests/test_kernels_parity.py -q -m "not gpu" -p no:cacheprovider ${KEXPR:+-k "$KEXPR"}
