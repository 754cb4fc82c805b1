#!/bin/bash
# usage: bash scripts/gpu_sweep_multi.sh N   (config-5 sample sweep, slab-sharded over N ranks, then the bench at N)
N=${1:-8}
mkdir -p gpurun_out; rm -f gpurun_out/rc_sweep.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --config 5 --size 512 > gpurun_out/sweep512_n$N.log 2>&1; echo "sweep n=$N rc=$?" >> gpurun_out/rc_sweep.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench n=$N rc=$?" >> gpurun_out/rc_sweep.txt
cat gpurun_out/rc_sweep.txt; grep "^{" gpurun_out/sweep512_n$N.log; tail -1 gpurun_out/bench_n$N.log | cut -c1-1100
