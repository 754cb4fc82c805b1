#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -x > gpurun_out/t_layers.log 2>&1; echo "layers rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/time_convs.py > gpurun_out/time_convs_rs1.log 2>&1; echo "rs1 rc=$?" >> gpurun_out/rc.txt
PMU_CONV_RS=0 timeout 300 python scripts/time_convs.py > gpurun_out/time_convs_rs0.log 2>&1; echo "rs0 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_layers.log; cat gpurun_out/time_convs_rs1.log; tail -1 gpurun_out/time_convs_rs0.log
