#!/bin/bash
# first GPU visit: unit parity, tcgen05 diagnostics, pipeline parity, smoke, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_gather_accum.py -m gpu -q -rA --no-header > gpurun_out/t_gather.log 2>&1; echo "gather rc=$?" >> gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q -rA --no-header -k "not conv_gemm and not fcomb_softmax" > gpurun_out/t_layers.log 2>&1; echo "layers rc=$?" >> gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q -rA --no-header -k "fcomb_softmax" > gpurun_out/t_fcomb.log 2>&1; echo "fcomb rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/diag_conv_tc.py > gpurun_out/diag_conv.log 2>&1; echo "diag rc=$?" >> gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q -rA --no-header -k "conv_gemm" > gpurun_out/t_convtc.log 2>&1; echo "convtc rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -rA --no-header > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -5 gpurun_out/bench.log
