set -u
mkdir -p gpurun_out/two
P=29541
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P scripts/dist_check.py --backend fused --reps 1 --height 1260 --width 1920 --loops 20 > gpurun_out/two/stream.txt 2>&1; grep -h "dist_check\|DIST_CHECK" gpurun_out/two/stream.txt | cut -c1-260
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) scripts/dist_check.py --backend fused --mode resident --reps 1 --height 1260 --width 1920 --loops 20 > gpurun_out/two/resident.txt 2>&1; grep -h "dist_check\|DIST_CHECK" gpurun_out/two/resident.txt | cut -c1-260
