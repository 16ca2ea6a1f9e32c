#!/bin/bash
mkdir -p gpurun_out/r2f
for cfg in "resident 16 5 14 0" "resident 16 5 14 8" "resident 16 5 14 16" "resident 16 5 14 24" "resident 16 5 14 1"; do
  set -- $cfg
  echo "== $1 W=$2 k=$3 R=$4 skip=$5" >> gpurun_out/r2f/skip.txt
  PCONV_PERSIST_SKIP=$5 timeout 120 python scripts/quick_bench.py --mode $1 --warps $2 --k $3 --band $4 >> gpurun_out/r2f/skip.txt 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2f/skip.txt'):
    if l.startswith('=='): print(l.strip(), end='  ')
    elif l.startswith('{'):
        d=json.loads(l); print('ms', d['ms'], {k: round(v,1) for k,v in d['probe'].items()})
    else: print(l.strip()[:200])
PY
