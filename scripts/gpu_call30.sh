#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_gather_accum.py -m gpu -q --no-header -rf -x > gpurun_out/t_gather.log 2>&1; echo "gather rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -5 gpurun_out/t_gather.log
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:(round(v['frac'],3), round(v['achieved'])) for k,v in d['hbm_kernels'].items()}, d['clocks'])"
