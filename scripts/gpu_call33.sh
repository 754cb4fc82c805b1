#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py -m gpu -q --no-header -rf -x > gpurun_out/t_layers.log 2>&1; echo "layers+model rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/time_convs.py > gpurun_out/time_convs_epg.log 2>&1; echo "convs rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -5 gpurun_out/t_layers.log; cat gpurun_out/time_convs_epg.log
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003}, d['roofline']['frac'], d['clocks'])"
