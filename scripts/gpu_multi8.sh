#!/bin/bash
# 8-GPU validation + timing (every GPU-minute counts 8x: keep it short)
N=8
OUT=gpurun_out/multi_$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "fused" > $OUT/pytest_multi.txt 2>&1; echo "pytest_multi rc=$?" >> $OUT/pytest_multi.txt; tail -3 $OUT/pytest_multi.txt
CASES=("--mode auto" "--mode resident --warps 16 --halo-depth 10" "--mode resident --warps 16 --halo-depth 15" "--mode resident --warps 8 --halo-depth 10" "--mode stream" "--mode resident --warps 16 --halo-depth 12 --grid 8x1")
for args in "${CASES[@]}"; do
  echo "== $args" >> $OUT/dist_headline.txt
  run scripts/dist_check.py --height 2520 --width 1920 --channels 3 --loops 100 --reps 3 $args 2>&1 | grep "dist_check\|DIST_CHECK\|Error\|error" | cut -c1-900 >> $OUT/dist_headline.txt
done
grep -c "DIST_CHECK OK" $OUT/dist_headline.txt; grep "DIST_CHECK FAILED\|MISMATCH\|rror" $OUT/dist_headline.txt | head -5
python - $OUT/dist_headline.txt <<'PY'
import re,sys
for l in open(sys.argv[1]):
    if l.startswith('=='): print(l.strip())
    m=re.search(r"rep=(\d) .*grid=(\S+) .*elapsed_ms=(\S+) max_wait_us=(\S+) .*mode=(\S+) plan=(\{.*?\}) probe=(\{.*?\})", l)
    if m and m.group(1)=='2': print('   ', m.group(2), 'ms', m.group(3), 'maxwait', m.group(4), m.group(5), m.group(6)[:120], m.group(7)[:260])
PY
run bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - $OUT/bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches','n_gpus')}, 'e2e', d['e2e']['ms_per_step'], d['config']['kernel_mode'], d['config']['plan'], d['config']['autotune'], d['config'].get('exchange_probe'), d['config'].get('useful_arithmetic_fraction'))
except Exception as e: print('bench parse failed', e)
PY
tail -3 $OUT/bench.err
