#!/bin/bash
mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests/test_gpu_resident.py -x -q > gpurun_out/r2d/pytest_resident.txt 2>&1
echo "pytest_resident rc=$?" >> gpurun_out/r2d/pytest_resident.txt
tail -4 gpurun_out/r2d/pytest_resident.txt
for cfg in "16 5 14 0" "16 5 14 7" "16 4 14 0" "8 3 14 0"; do
  set -- $cfg
  echo "== W=$1 k=$2 R=$3 skip=$4" >> gpurun_out/r2d/skip.txt
  PCONV_PERSIST_SKIP=$4 timeout 120 python scripts/quick_bench.py --mode resident --warps $1 --k $2 --band $3 >> gpurun_out/r2d/skip.txt 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2d/skip.txt'):
    if l.startswith('=='): print(l.strip(), end='  ')
    elif l.startswith('{'):
        d=json.loads(l); print('ms', d['ms'], {k: round(v,1) for k,v in d['probe'].items()})
    else: print(l.strip()[:200])
PY
python -c "
import numpy as np
np.random.default_rng(1).integers(0,256,1920*2520*3,dtype=np.uint8).tofile('/tmp/img.raw')"
cd /tmp && for i in 1 2 3; do PCONV_NATIVE_TIMING=1 $GRAFT_REPO_ROOT/parallel_convolution_b200/_lib/pconv_native img.raw 1920 2520 100 rgb; done 2>&1 | tail -24
for i in 1 2; do /usr/bin/time -f "ref wall %e s" $GRAFT_REPO_ROOT/baseline/_ref/cuda/cuda_conv img.raw 1920 2520 100 rgb; done 2>&1 | tail -4
