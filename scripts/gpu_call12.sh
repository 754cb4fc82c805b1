#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/time_convs.py 64 > gpurun_out/time_convs.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -3 gpurun_out/t_all.log; grep -E "\.T |TOTAL" gpurun_out/time_convs.log
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003})"
