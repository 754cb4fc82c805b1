#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -rf > gpurun_out/t_train.log 2>&1; echo "train rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -60 gpurun_out/t_train.log
