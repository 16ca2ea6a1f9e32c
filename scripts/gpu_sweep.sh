#!/bin/bash
# usage: gpu_sweep.sh N [edges] — halo sweep (k=1 and planner k) + headline bench of both backends on N GPUs
N=${1:-2}
EDGES=${2:-256,1024,4096}
OUT=gpurun_out/multi_$N
mkdir -p $OUT
run() { timeout ${TMO:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
run bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - $OUT/bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches','n_gpus')}, 'e2e', d['e2e']['ms_per_step'], d['config']['kernel_mode'], d['config']['plan'], d['config']['autotune'], d['config'].get('exchange_probe'))
except Exception as e: print('bench parse failed', e)
PY
TMO=120 run bench.py --gpus $N --steps 20 --warmup 5 --backend nccl > $OUT/bench_nccl.json 2> $OUT/bench_nccl.err; echo "bench nccl(eager) rc=$?"
python - $OUT/bench_nccl.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('NCCL arm:', {k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches','n_gpus')}, 'e2e', d['e2e']['ms_per_step'], d['config']['plan'])
except Exception as e: print('nccl bench parse failed', e)
PY
tail -2 $OUT/bench_nccl.err | cut -c1-300
TMO=240 run bench/halo_sweep.py --edges $EDGES --loops 40 --halo-depth 1 --with-cpu > $OUT/halo_sweep_k1.jsonl 2> $OUT/halo_sweep_k1.err; echo "sweep k1 rc=$?"
TMO=240 run bench/halo_sweep.py --edges $EDGES --loops 40 --halo-depth 0 > $OUT/halo_sweep_kplan.jsonl 2> $OUT/halo_sweep_kplan.err; echo "sweep kplan rc=$?"
python - $OUT <<'PY'
import json,sys
for f in ('halo_sweep_k1.jsonl','halo_sweep_kplan.jsonl'):
    print(f)
    for l in open(sys.argv[1]+'/'+f):
        if l.startswith('{'):
            d=json.loads(l); print('  ', {k:v for k,v in d.items() if k.endswith('_us_per_iter') or k in('tile_edge','grid') or k.endswith('_k') or k.endswith('_skipped')})
PY
tail -3 $OUT/halo_sweep_k1.err | cut -c1-300
