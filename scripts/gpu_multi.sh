#!/bin/bash
# 8-GPU round check (gpurun --gpus 8): headline bench with the e2e phase timings, the broadcast-upload / graph variants of the
# e2e leg, BASELINE config 4 (DP training step) and config 5 (512^3 sample sweep, slab-sharded)
N=${1:-8}
mkdir -p gpurun_out; rm -f gpurun_out/multi_*.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T bench.py --gpus $N --steps 10 --warmup 3 --e2e-phases > gpurun_out/multi_bench.log 2>&1; echo "bench rc=$?"
timeout 400 $T bench.py --gpus $N --steps 10 --warmup 3 --e2e-upload broadcast > gpurun_out/multi_bench_bcast.log 2>&1; echo "bcast rc=$?"
timeout 400 $T bench.py --gpus $N --steps 10 --warmup 3 --graph > gpurun_out/multi_bench_graph.log 2>&1; echo "graph rc=$?"
timeout 400 $T bench.py --gpus $N --config 4 --steps 8 > gpurun_out/multi_cfg4.log 2>&1; echo "cfg4 rc=$?"
timeout 600 $T bench.py --gpus $N --config 5 --sweep-samples 1,16,128 --sweep-steps 1 > gpurun_out/multi_cfg5.log 2>&1; echo "cfg5 rc=$?"
for f in bench bench_bcast bench_graph; do tail -1 gpurun_out/multi_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$f', round(d['value'],2), 'vol/s resident', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],2), 'ms', d['e2e'].get('phases_ms'))
except Exception as ex: print('$f parse error', ex)"; done
tail -1 gpurun_out/multi_cfg4.log | cut -c1-600; grep '^{' gpurun_out/multi_cfg5.log | cut -c1-330
