#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -x -k "conv_gemm" > gpurun_out/t_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
for v in 0 1 2; do
  PMU_CONV_VARIANT=$v timeout 300 python scripts/time_convs.py 64 > gpurun_out/time_convs_v$v.log 2>&1; echo "v$v rc=$?" >> gpurun_out/rc.txt
done
PMU_CONV_VARIANT=2 timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -x -k "conv_gemm" > gpurun_out/t_conv_v2.log 2>&1; echo "pytest v2 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -3 gpurun_out/t_conv.log; tail -3 gpurun_out/t_conv_v2.log
