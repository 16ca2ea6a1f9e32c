#!/bin/bash
# One-box validation of the tree as committed: build + smoke, the GPU test suite, the headline bench
# (both arms), the design microbenchmarks, a launch-overhead fit and the sanitizer subset. Everything
# lands in gpurun_out/final/ (copy what should be kept into profiles/).
set -u
O=gpurun_out/final
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)"
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)"
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? $(cut -c1-300 $O/bench_n1.json)"
timeout 60 scripts/microbench/f32x2 > $O/f32x2.txt 2>&1; cat $O/f32x2.txt
timeout 60 scripts/microbench/flag_latency > $O/flag_latency.txt 2>&1; cat $O/flag_latency.txt
for k in 1 2 3 4 5; do timeout 120 python scripts/quick_bench.py --k $k --band 14 --loops 60 --reps 5; done > $O/kfit.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/final/kfit.txt'):
    try: d=json.loads(l)
    except Exception: continue
    p=d['plan']; n=-(-d['loops']//p['k'])
    print('kfit k=%d R=%d ctas=%d launches=%d ms=%.4f us/launch=%.2f' % (p['k'],p['R'],p['ntx']*p['nty'],n,d['ms'],d['ms']*1e3/n))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$? $(cut -c1-200 $O/bench_ref.json)"
bash scripts/sanitize.sh > $O/sanitize.txt 2>&1; cat $O/sanitize.txt
