#!/bin/bash
mkdir -p gpurun_out/ncu_r2
ncu --set full --clock-control none --import-source on -k regex:persist_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r2/persist_w16_k5_R14 \
    python scripts/quick_bench.py --mode resident --warps 16 --k 5 --band 14 --reps 2 > gpurun_out/ncu_r2/persist_run.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2/persist_run.log; ls -la gpurun_out/ncu_r2/
