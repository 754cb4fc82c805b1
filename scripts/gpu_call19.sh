#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 600 python scripts/bench_train.py 8 3 --cpu > gpurun_out/bench_train.log 2>&1; echo "train rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_all.log; cat gpurun_out/bench_train.log
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003}, {k:round(v['frac'],3) for k,v in d['hbm_kernels'].items()}, d['roofline']['frac'], d['cpu_baseline'])"
