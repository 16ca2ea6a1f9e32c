#!/bin/bash
# final single-GPU evidence: full GPU test-suite, bench both arms, ncu of the default kernel, sanitizer subset
mkdir -p gpurun_out/final_r2
O=gpurun_out/final_r2
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ('bench_n1','bench_ref'):
    try:
        d=json.loads(open(f'gpurun_out/final_r2/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches')}, d.get('e2e',{}).get('value'), d.get('whole_program'), d['config'].get('kernel_mode'), d['config'].get('plan'), d['config'].get('rebuilt_sm_100a'))
    except Exception as e: print(f, 'parse failed', e)
PY
ncu --set full --clock-control none --import-source on -k regex:persist_kernel -s 2 -c 1 -f -o $O/persist_w16_k5_R14 \
    python scripts/quick_bench.py --mode resident --warps 16 --k 5 --band 14 --reps 2 > $O/ncu_persist.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-autotune > $O/launches_bench.log 2>&1
TOOLS="memcheck synccheck" SAN_TIMEOUT=300 SAN_FILTER="small_tiles or 3-2-None or 97-333" bash scripts/sanitize.sh 2>&1 | tail -3
cp gpurun_out/sanitizer_*.log $O/ 2>/dev/null
