#!/bin/bash
# one-GPU round check (run under gpurun): parity suite, smoke, bench (ours + reference arm), training-step timing
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?" >> gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 600 python tests/tools/bench_train.py 8 8 --bf16 > gpurun_out/bench_train_bf16.log 2>&1; echo "train bf16 rc=$?" >> gpurun_out/rc.txt
timeout 600 python tests/tools/bench_train.py 8 4 > gpurun_out/bench_train_fp32.log 2>&1; echo "train fp32 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -4 gpurun_out/t_all.log; tail -3 gpurun_out/smoke.log; grep "^train step\|^per-step" gpurun_out/bench_train_*.log
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],1) for k,v in d['kernel_time_shares'].items() if v>0.003}, {k:round(v['frac'],3) for k,v in d['hbm_kernels'].items()}, d['roofline']['frac'], d['cpu_baseline'], d['clocks'])"
tail -1 gpurun_out/bench_ref.log | cut -c1-300
