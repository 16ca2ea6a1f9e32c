#!/bin/bash
# attribution of the exchange cost: skip phases one at a time (results wrong, timing only)
mkdir -p gpurun_out/r2b
for cfg in "16 5 14" "8 3 14"; do
  set -- $cfg
  for skip in 0 1 2 4 3 6 7; do
    echo "== W=$1 k=$2 R=$3 skip=$skip" >> gpurun_out/r2b/skip.txt
    PCONV_PERSIST_SKIP=$skip timeout 120 python scripts/quick_bench.py --mode resident --warps $1 --k $2 --band $3 >> gpurun_out/r2b/skip.txt 2>&1
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2b/skip.txt'):
    if l.startswith('=='): print(l.strip(), end='  ')
    elif l.startswith('{'):
        d=json.loads(l); print('ms', d['ms'], {k: round(v,1) for k,v in d['probe'].items()})
PY
