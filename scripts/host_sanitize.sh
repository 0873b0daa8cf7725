#!/bin/bash
# ASan + UBSan over the HOST code of libcustos_b200 (expression IR, code generator, NVRTC driver, OptGraph,
# serde codec, argument checking): builds the instrumented library and runs the CPU test-suite against it.
# Leak checking is off: the interpreter itself never frees everything.
set -e
cd "$(dirname "$0")/.."
python -m custos_b200.build
python -m custos_b200.build --sanitize
make -s -C oracle
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so)
LD_PRELOAD="$ASAN:$UBSAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  CUSTOS_B200_LIB=custos_b200/lib/libcustos_b200_asan.so \
  python -m pytest tests -q -m "not gpu" -p no:cacheprovider "$@"
