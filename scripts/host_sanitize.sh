#!/bin/bash
# ASan + UBSan over the native host library (oracle, CPU engine, raw I/O, decomposition, planner):
# builds an instrumented libpconv_host_asan.so and runs the CPU test-suite against it.
set -e
cd "$(dirname "$0")/.."
LIBASAN=$(gcc -print-file-name=libasan.so)
python -c "from parallel_convolution_b200 import build; print(build.build_host_sanitized())"
ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1 \
LD_PRELOAD=$LIBASAN PCONV_HOST_LIB=$PWD/parallel_convolution_b200/_lib/libpconv_host_asan.so \
  python -m pytest tests/test_cpu_engine.py tests/test_engine_cpu.py tests/test_xplan_emu.py -q -x -p no:cacheprovider tests/test_shm_ranks.py

# ThreadSanitizer over the shared-image CPU ranks: the ranks of tests/test_shm_ranks.py are threads of one
# process, so TSan sees every image access of every rank and checks that the epoch counters (acquire/release)
# order them. Ranks run single-threaded here: libgomp is not instrumented and its barriers would be reported.
LIBTSAN=$(gcc -print-file-name=libtsan.so)
CS=parallel_convolution_b200/csrc/host
g++ -std=c++17 -fPIC -O1 -g -ffp-contract=off -fopenmp -fno-omit-frame-pointer -fsanitize=thread -Wno-tsan -shared \
    -o parallel_convolution_b200/_lib/libpconv_host_tsan.so $CS/filters.cpp $CS/decomp.cpp $CS/cpu_engine.cpp \
    $CS/rawio.cpp $CS/planner.cpp $CS/xplan_emu.cpp $CS/shm_ranks.cpp $CS/oracle.cpp
TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0 exitcode=66" PCONV_TEST_SINGLE_THREAD_RANKS=1 \
LD_PRELOAD=$LIBTSAN PCONV_HOST_LIB=$PWD/parallel_convolution_b200/_lib/libpconv_host_tsan.so \
  python -m pytest tests/test_shm_ranks.py -q -x -p no:cacheprovider -k "thread_ranks or reload"
