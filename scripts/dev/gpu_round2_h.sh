#!/bin/bash
mkdir -p gpurun_out/r2h
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r2h/clocks_before.txt
timeout 120 scripts/microbench/issue_peak > gpurun_out/r2h/issue_peak.txt 2>&1; echo "issue_peak rc=$?"; cat gpurun_out/r2h/issue_peak.txt
timeout 60 scripts/microbench/tc_toeplitz > gpurun_out/r2h/tc_toeplitz.txt 2>&1; echo "tc rc=$?"; cat gpurun_out/r2h/tc_toeplitz.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h/pytest_gpu.txt; tail -4 gpurun_out/r2h/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h/bench1.json 2> gpurun_out/r2h/bench1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h/bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','verified','gpu_launches')}, d['e2e']['ms_per_step'], d['config']['kernel_mode'], d['config']['plan'], d['config']['autotune'])
PY
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
