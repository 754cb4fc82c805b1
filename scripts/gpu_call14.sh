#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --interp exact > gpurun_out/bench_exact.log 2>&1; echo "bench exact rc=$?" >> gpurun_out/rc.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_all.log
for f in bench bench_exact; do tail -1 gpurun_out/$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003}, {k:round(v['frac'],3) for k,v in d['hbm_kernels'].items()}, d['roofline']['traffic'], d['cpu_baseline'])"; done
tail -1 gpurun_out/bench_ref.log | cut -c1-300
