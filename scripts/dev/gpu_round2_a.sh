#!/bin/bash
# first contact of the persistent kernel with hardware: numerics, then a mode comparison on the headline image
mkdir -p gpurun_out/r2a
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi -L > gpurun_out/r2a/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_resident.py -x -q > gpurun_out/r2a/pytest_resident.txt 2>&1
echo "pytest_resident rc=$?" >> gpurun_out/r2a/pytest_resident.txt
tail -5 gpurun_out/r2a/pytest_resident.txt
for cfg in "stream 0 0 0" "resident 8 3 14" "resident 8 2 14" "resident 16 5 14" "resident 16 3 14" "resident 8 3 12" "resident 8 4 14"; do
  set -- $cfg
  timeout 120 python scripts/quick_bench.py --mode $1 --warps $2 --k $3 --band $4 >> gpurun_out/r2a/modes.txt 2>&1
done
cat gpurun_out/r2a/modes.txt | cut -c1-400
