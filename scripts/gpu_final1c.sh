#!/bin/bash
# validation of a kernel change on one GPU within a tight budget: full GPU test-suite, then the headline bench
O=gpurun_out/final_r2
mkdir -p $O
timeout 260 python -m pytest tests -x -q -m gpu > $O/pytest_gpu_c.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_c.txt; tail -3 $O/pytest_gpu_c.txt
timeout 40 python scripts/cta_profile.py > $O/cta_profile_c.txt 2>&1; tail -14 $O/cta_profile_c.txt | head -13
timeout 120 python bench.py --steps 20 --warmup 5 > $O/bench_n1_c.json 2> $O/bench_n1_c.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/final_r2/bench_n1_c.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches')}, d['e2e']['ms_per_step'], d['config']['kernel_mode'], d['config']['plan'], d['config']['autotune'])
except Exception as e: print('parse failed', e)
PY
