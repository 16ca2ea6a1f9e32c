#!/bin/bash
# final 8-GPU evidence (each minute is charged 8x): correctness of the default mode on the headline job,
# headline bench of both backends, a short halo sweep
N=8
OUT=gpurun_out/multi_$N
mkdir -p $OUT
run() { timeout ${TMO:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
rm -f $OUT/dist_final.txt
for args in "--mode auto" "--mode stream"; do
  echo "== $args" >> $OUT/dist_final.txt
  TMO=90 run scripts/dist_check.py --height 2520 --width 1920 --channels 3 --loops 100 --reps 3 $args 2>&1 | grep "dist_check\|DIST_CHECK" | cut -c1-700 >> $OUT/dist_final.txt
done
grep "rep=2\|DIST_CHECK" $OUT/dist_final.txt | cut -c1-330
TMO=150 run bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
TMO=100 run bench.py --gpus $N --steps 20 --warmup 5 --backend nccl > $OUT/bench_nccl.json 2> $OUT/bench_nccl.err; echo "bench nccl rc=$?"
python - $OUT <<'PY'
import json,sys
for f in ('bench.json','bench_nccl.json'):
    try:
        d=json.loads(open(sys.argv[1]+'/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches','n_gpus')}, 'e2e', d['e2e']['ms_per_step'], d['config']['kernel_mode'], d['config']['plan'], d['config'].get('autotune'), d['config'].get('exchange_probe'))
    except Exception as e: print(f, 'parse failed', e)
PY
TMO=150 run bench/halo_sweep.py --edges 256,1024 --loops 40 --halo-depth 1 > $OUT/halo_sweep_k1.jsonl 2> $OUT/halo_sweep_k1.err; echo "sweep k1 rc=$?"
TMO=150 run bench/halo_sweep.py --edges 256,1024,4096 --loops 40 --halo-depth 0 > $OUT/halo_sweep_kplan.jsonl 2> $OUT/halo_sweep_kplan.err; echo "sweep kplan rc=$?"
python - $OUT <<'PY'
import json,sys
for f in ('halo_sweep_k1.jsonl','halo_sweep_kplan.jsonl'):
    print(f)
    for l in open(sys.argv[1]+'/'+f):
        if l.startswith('{'):
            d=json.loads(l); print('  ', {k:v for k,v in d.items() if k.endswith('_us_per_iter') or k in('tile_edge','grid') or k.endswith('_k')})
PY
