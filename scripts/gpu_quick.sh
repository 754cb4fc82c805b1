#!/bin/bash
# quick one-GPU check: whole parity suite, smoke, default bench
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf --durations=8 > gpurun_out/t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -16 gpurun_out/t_all.log; tail -3 gpurun_out/smoke.log
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003}, d['roofline']['frac'], d['cpu_baseline'], d['clocks'])"
