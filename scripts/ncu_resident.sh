set -u
mkdir -p gpurun_out/ncu_res
timeout 500 ncu --set full --clock-control none --import-source on -k regex:resident_kernel -c 1 -o gpurun_out/ncu_res/res14 \
    python scripts/quick_bench.py --mode resident --loops 100 --reps 1 > gpurun_out/ncu_res/log.txt 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_res/log.txt | cut -c1-200
ncu -i gpurun_out/ncu_res/res14.ncu-rep --page source --csv > gpurun_out/ncu_res/res14_source.csv 2>gpurun_out/ncu_res/src_err.txt
ncu -i gpurun_out/ncu_res/res14.ncu-rep --page raw --csv > gpurun_out/ncu_res/res14_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_res/
