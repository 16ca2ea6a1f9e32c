#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests/test_gpu_resident.py -x -q > gpurun_out/r2c/pytest_resident.txt 2>&1
echo "pytest_resident rc=$?" >> gpurun_out/r2c/pytest_resident.txt
tail -4 gpurun_out/r2c/pytest_resident.txt
for cfg in "16 5 14 0" "16 5 14 7" "8 3 14 0"; do
  set -- $cfg
  echo "== W=$1 k=$2 R=$3 skip=$4" >> gpurun_out/r2c/skip.txt
  PCONV_PERSIST_SKIP=$4 timeout 120 python scripts/quick_bench.py --mode resident --warps $1 --k $2 --band $3 >> gpurun_out/r2c/skip.txt 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c/skip.txt'):
    if l.startswith('=='): print(l.strip(), end='  ')
    elif l.startswith('{'):
        d=json.loads(l); print('ms', d['ms'], {k: round(v,1) for k,v in d['probe'].items()})
    else: print(l.strip()[:200])
PY
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c/bench1.json 2> gpurun_out/r2c/bench1.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2c/bench1.json; tail -5 gpurun_out/r2c/bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c/ref1.json 2> gpurun_out/r2c/ref1.err
echo "ref rc=$?"; tail -c 2500 gpurun_out/r2c/ref1.json; tail -5 gpurun_out/r2c/ref1.err
