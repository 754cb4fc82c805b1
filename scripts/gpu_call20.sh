#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -rf -k "bf16" > gpurun_out/t_wgrad.log 2>&1; echo "wgrad rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -40 gpurun_out/t_wgrad.log
