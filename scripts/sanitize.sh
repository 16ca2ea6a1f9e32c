#!/bin/bash
# Memory-safety / race / sync checks of the sm_100a kernels under compute-sanitizer (run on a GPU
# box: `gpurun -- bash scripts/sanitize.sh`). The reference's RGB kernel reads and writes out of
# bounds (defect D3) and would fail the first of these. Output goes to gpurun_out/sanitizer_*.log;
# a one-line-per-tool summary is what gets committed under profiles/.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
TESTS="tests/test_gpu_stream.py::test_filters_match_oracle tests/test_gpu_stream.py::test_remainder_tiles tests/test_gpu_stream.py::test_small_tiles_in_deep_grids tests/test_gpu_resident.py::test_resident_virtual_ranks tests/test_gpu_resident.py::test_resident_odd_shapes"
TOOLS=${TOOLS:-"memcheck racecheck synccheck"}
for tool in $TOOLS; do
  timeout ${SAN_TIMEOUT:-1500} $CS --tool $tool --error-exitcode 9 --launch-timeout 120 \
      python -m pytest $TESTS -m gpu -x -q -k "${SAN_FILTER:-gaussian or sharpen or box or remainder or small_tiles or 2-None or grid2 or 211-997 or 97-333}" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
