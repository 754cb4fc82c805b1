#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -x -k first_conv > gpurun_out/t_first.log 2>&1; echo "first rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --no-header -rf -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -12 gpurun_out/t_first.log; tail -4 gpurun_out/t_model.log
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003}, d['clocks'])"
