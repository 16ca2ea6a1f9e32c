#!/bin/bash
mkdir -p gpurun_out/r2g
timeout 1200 python -m pytest tests/test_gpu_resident.py tests/test_gpu_stream.py tests/test_gpu_vs_torch.py -x -q > gpurun_out/r2g/pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g/pytest.txt
tail -4 gpurun_out/r2g/pytest.txt
python scripts/cta_profile.py > gpurun_out/r2g/cta_profile.txt 2>&1; tail -16 gpurun_out/r2g/cta_profile.txt
for cfg in "resident 16 5 14 0" "resident 16 5 14 7" "resident 8 3 14 0" "stream 0 0 0 0"; do
  set -- $cfg
  echo "== $1 W=$2 k=$3 R=$4 skip=$5" >> gpurun_out/r2g/skip.txt
  PCONV_PERSIST_SKIP=$5 timeout 120 python scripts/quick_bench.py --mode $1 --warps $2 --k $3 --band $4 >> gpurun_out/r2g/skip.txt 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2g/skip.txt'):
    if l.startswith('=='): print(l.strip(), end='  ')
    elif l.startswith('{'):
        d=json.loads(l); print('ms', d['ms'], {k: round(v,1) for k,v in d['probe'].items()})
    else: print(l.strip()[:200])
PY
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g/bench1.json 2> gpurun_out/r2g/bench1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g/bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','verified','gpu_launches')}, d['e2e']['ms_per_step'], d['config']['kernel_mode'], d['config']['plan'], d['config']['autotune'], d['config'].get('exchange_probe'), d.get('whole_program'))
PY
tail -3 gpurun_out/r2g/bench1.err
