#!/bin/bash
# usage: bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_n$N.log 2>&1; echo "bench n=$N rc=$?" >> gpurun_out/rc_multi.txt
cat gpurun_out/rc_multi.txt; tail -3 gpurun_out/bench_n$N.log
