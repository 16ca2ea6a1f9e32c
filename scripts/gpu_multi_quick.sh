#!/bin/bash
# usage: gpu_multi_quick.sh N — real multi-GPU tests + headline bench (fused) on N GPUs
N=${1:-2}
OUT=gpurun_out/multi_$N
mkdir -p $OUT
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $OUT/pytest_multi.txt 2>&1; echo "pytest_multi rc=$?" >> $OUT/pytest_multi.txt; tail -3 $OUT/pytest_multi.txt
for args in "--mode stream" "--mode auto"; do
  echo "== $args" >> $OUT/dist_quick.txt
  run scripts/dist_check.py --height 2520 --width 1920 --channels 3 --loops 100 --reps 3 $args 2>&1 | grep "dist_check\|DIST_CHECK" | cut -c1-400 >> $OUT/dist_quick.txt
done
grep "rep=2\|DIST_CHECK" $OUT/dist_quick.txt | cut -c1-260
run bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - $OUT/bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches','n_gpus')}, 'e2e', d['e2e']['ms_per_step'], d['config']['kernel_mode'], d['config']['plan'], d['config']['autotune'], d['config'].get('exchange_probe'))
except Exception as e: print('bench parse failed', e)
PY
