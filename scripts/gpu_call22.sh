#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -rf -s -k "bf16_training" > gpurun_out/t_train16.log 2>&1; echo "rc=$?"
tail -30 gpurun_out/t_train16.log
