#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/diag_conv_tc.py > gpurun_out/diag_conv.log 2>&1; echo "diag rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_all.log; tail -6 gpurun_out/diag_conv.log; tail -1 gpurun_out/bench.log
