#!/bin/bash
# last check of the committed tree on one GPU: tuner-dependent tests, headline bench, BASELINE configs 3 and 4 on one GPU
O=gpurun_out/final_r2
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_resident.py -x -q -k "autotune or single_rank or back_to_back or falls_back" > $O/pytest_final_subset.txt 2>&1; tail -2 $O/pytest_final_subset.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_n1_final.json 2> $O/bench_n1_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_r2/bench_n1_final.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches')}, d['e2e']['ms_per_step'], d['config']['kernel_mode'], d['config']['plan'], d['config']['autotune'], d['config']['fraction_of_binding_roofline'])
PY
timeout 200 python bench/run_configs.py --config 3 --reps 1 > $O/config3_n1.json 2>&1; tail -1 $O/config3_n1.json | cut -c1-400
timeout 200 python bench/run_configs.py --config 4 --reps 1 > $O/config4_n1.json 2>&1; tail -1 $O/config4_n1.json | cut -c1-400
