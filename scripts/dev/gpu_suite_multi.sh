#!/bin/bash
# Multi-GPU measurement suite (run under `gpurun --gpus N -- bash scripts/gpu_suite_multi.sh N`).
# Everything lands in gpurun_out/multi_N/ ; summaries are copied to profiles/ by hand.
N=${1:-8}
OUT=gpurun_out/multi_$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpus.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $OUT/pytest_multi.log 2>&1; tail -2 $OUT/pytest_multi.log
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  timeout 300 python bench.py --gpus $n --steps 150 --warmup 10 2>/dev/null | grep '^{' > $OUT/bench_n$n.json
  python -c "import json;d=json.load(open('$OUT/bench_n$n.json'));print('bench n=$n', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['config']['plan'])"
done
timeout 300 python bench.py --gpus $N --steps 100 --warmup 10 --mode resident 2>/dev/null | grep '^{' > $OUT/bench_n${N}_resident.json
timeout 300 python bench.py --gpus $N --steps 100 --warmup 10 --backend nccl 2>/dev/null | grep '^{' > $OUT/bench_n${N}_nccl.json
for f in resident nccl; do python -c "import json;d=json.load(open('$OUT/bench_n${N}_$f.json'));print('bench $f n=$N', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"; done
timeout 600 $TR --master-port 29601 bench/run_configs.py --config 2 2>/dev/null | grep '^{' | tee $OUT/config2.json | cut -c1-400
timeout 900 $TR --master-port 29602 bench/run_configs.py --config 3 2>/dev/null | grep '^{' | tee $OUT/config3.json | cut -c1-400
timeout 900 $TR --master-port 29603 bench/run_configs.py --config 4 2>/dev/null | grep '^{' | tee $OUT/config4.json | cut -c1-400
timeout 900 $TR --master-port 29604 bench/run_configs.py --config 4 --backend nccl --reps 1 2>/dev/null | grep '^{' | tee $OUT/config4_nccl.json | cut -c1-400
timeout 900 $TR --master-port 29605 bench/halo_sweep.py --edges 256,1024,4096,16384 --loops 30 2>/dev/null | grep '^{' | tee $OUT/halo_sweep.json | cut -c1-500
