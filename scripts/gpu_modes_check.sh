# One-GPU check used during kernel work: full GPU test suite, the three kernel modes on 1 and 8 virtual
# ranks (quick_bench), the per-launch cost fit, and the headline bench. Output: gpurun_out/modes/.
set -u
O=gpurun_out/modes; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)"
for args in "--mode stream" "--mode resident" "--mode resident2" "--mode stream --workers 8" "--mode resident --workers 8" "--mode resident2 --workers 8"; do
  timeout 120 python scripts/quick_bench.py --loops 100 --reps 6 $args 2>&1 | tail -1 | cut -c1-700
done > $O/quick.txt 2>&1
cat $O/quick.txt
for k in 1 3 5; do timeout 120 python scripts/quick_bench.py --k $k --band 14 --loops 60 --reps 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['plan']; n=-(-d['loops']//p['k'])
print('kfit k=%d us/launch=%.2f' % (p['k'], d['ms']*1e3/n))"; done
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-420 $O/bench_n1.json
