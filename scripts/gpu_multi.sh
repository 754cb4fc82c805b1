#!/bin/bash
# usage: bash scripts/gpu_multi.sh N   (bench + reference arm under torchrun, DP training step on N ranks)
N=${1:-2}
mkdir -p gpurun_out; rm -f gpurun_out/rc_multi.txt
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench n=$N rc=$?" >> gpurun_out/rc_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref n=$N rc=$?" >> gpurun_out/rc_multi.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tests/tools/bench_train_dp.py > gpurun_out/train_dp_n$N.log 2>&1; echo "train dp n=$N rc=$?" >> gpurun_out/rc_multi.txt
cat gpurun_out/rc_multi.txt; tail -1 gpurun_out/bench_n$N.log | cut -c1-900; tail -2 gpurun_out/bench_ref_n$N.log | cut -c1-300; grep -v Warn gpurun_out/train_dp_n$N.log | tail -4
