#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_gather_accum.py -m gpu -q --no-header -rf > gpurun_out/t_gather.log 2>&1; echo "gather rc=$?" >> gpurun_out/rc.txt
for sb in 64 128 256; do
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --slice-batch $sb > gpurun_out/bench_sb$sb.log 2>&1; echo "bench sb=$sb rc=$?" >> gpurun_out/rc.txt
tail -1 gpurun_out/bench_sb$sb.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($sb, d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003}, {k:round(v['frac'],3) for k,v in d['hbm_kernels'].items()})"
done
cat gpurun_out/rc.txt; tail -3 gpurun_out/t_gather.log
