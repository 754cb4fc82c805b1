#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -k "fcomb_softmax or conv_gemm" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/run_fcomb.py 64 16 > gpurun_out/fcomb_plain.log 2>&1; echo "fcomb rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -3 gpurun_out/t_tc.log; cat gpurun_out/fcomb_plain.log; tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003})"
