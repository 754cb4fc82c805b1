#!/bin/bash
# training path check: parity tests of the step (+ the layer tests its kernels share), then the batch-8 step timing
mkdir -p gpurun_out; rm -f gpurun_out/rc_train.txt
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_layers.py -m gpu -q --no-header -rf > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?" >> gpurun_out/rc_train.txt
timeout 300 python tests/tools/prof_train_host.py 8 6 --no-profiler > gpurun_out/train_step_phases.log 2>&1; echo "phases rc=$?" >> gpurun_out/rc_train.txt
timeout 600 python tests/tools/bench_train.py 8 6 --bf16 > gpurun_out/bench_train_bf16.log 2>&1; echo "train bf16 rc=$?" >> gpurun_out/rc_train.txt
cat gpurun_out/rc_train.txt; tail -6 gpurun_out/t_train.log; grep "^it \|alloc" gpurun_out/train_step_phases.log; grep -A13 "^train step" gpurun_out/bench_train_bf16.log
