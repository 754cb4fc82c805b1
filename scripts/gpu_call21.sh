#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -m gpu -q --no-header -rf -s -k "bf16_training or unsupported or training_step" > gpurun_out/t_train16.log 2>&1; echo "t rc=$?" >> gpurun_out/rc.txt
timeout 600 python scripts/bench_train.py 8 3 --bf16 > gpurun_out/bench_train_bf16.log 2>&1; echo "train bf16 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -30 gpurun_out/t_train16.log; cat gpurun_out/bench_train_bf16.log
