#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -k "fcomb_softmax" > gpurun_out/t_fcomb.log 2>&1; echo "fcomb rc=$?" >> gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -k "conv_gemm" > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?" >> gpurun_out/rc.txt
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/time_convs.py 64 > gpurun_out/time_convs.log 2>&1; echo "time rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
PMU_FCOMB_MMA_SYNC=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_mma.log 2>&1; echo "bench_mma rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -12 gpurun_out/t_fcomb.log; tail -4 gpurun_out/t_all.log; tail -3 gpurun_out/time_convs.log; tail -1 gpurun_out/bench.log
