#!/bin/bash
# ASan + UBSan over the native host library (oracle, CPU engine, raw I/O, decomposition, planner):
# builds an instrumented libpconv_host_asan.so and runs the CPU test-suite against it.
set -e
cd "$(dirname "$0")/.."
LIBASAN=$(gcc -print-file-name=libasan.so)
python -c "from parallel_convolution_b200 import build; print(build.build_host_sanitized())"
ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1 \
LD_PRELOAD=$LIBASAN PCONV_HOST_LIB=$PWD/parallel_convolution_b200/_lib/libpconv_host_asan.so \
  python -m pytest tests/test_cpu_engine.py tests/test_engine_cpu.py tests/test_xplan_emu.py -q -x -p no:cacheprovider
