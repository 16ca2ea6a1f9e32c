# First hardware contact of the protocol-warp kernel: correctness (opt-in tests, bounded by timeout: a
# named-barrier mistake would hang), then the four kernel modes on a half-height image.
set -u
mkdir -p gpurun_out/r3
PCONV_TEST_RESIDENT3=1 timeout 50 python -m pytest tests/test_gpu_resident3.py -x -q > gpurun_out/r3/pytest.log 2>&1; echo "pytest rc=$? $(tail -2 gpurun_out/r3/pytest.log | tr '\n' ' ')"
for m in stream resident resident2 resident3; do
  timeout 25 python scripts/quick_bench.py --height 1260 --loops 100 --reps 5 --mode $m 2>&1 | tail -1 | cut -c1-420
done > gpurun_out/r3/quick.txt 2>&1
cat gpurun_out/r3/quick.txt
