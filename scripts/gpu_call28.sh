#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --no-header -rf > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -5 gpurun_out/t_model.log
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['clocks'])"
