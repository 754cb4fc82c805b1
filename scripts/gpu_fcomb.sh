#!/bin/bash
# fcomb kernel variants: parity, then the kernel alone (64 slices of 256x256, 16 samples) per variant, then a timeline of ts2
mkdir -p gpurun_out; rm -f gpurun_out/fc_*.log
timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -x -k "fcomb_softmax_accum_bf16 and ts2" > gpurun_out/fc_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/fc_tests.log
tail -5 gpurun_out/fc_tests.log
for v in 0 3; do
  PMU_FCOMB_TS=$v timeout 120 python scripts/run_fcomb.py 64 16 2>&1 | tail -1 | sed "s/^/TS=$v /" | tee -a gpurun_out/fc_time.log
done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I probabilistic-multiplanar-unet_b200/csrc -o /tmp/fcomb_trace scripts/fcomb_trace.cu 2>/dev/null && timeout 60 /tmp/fcomb_trace 32 > gpurun_out/fcomb_trace.log; echo trace rc=$?; head -3 gpurun_out/fcomb_trace.log
