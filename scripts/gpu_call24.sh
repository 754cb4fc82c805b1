#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
PMU_FCOMB_TS=1 timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -x -k fcomb_softmax > gpurun_out/t_fcomb.log 2>&1; echo "fcomb rc=$?" >> gpurun_out/rc.txt
PMU_FCOMB_TS=1 timeout 300 python scripts/run_fcomb.py 64 16 > gpurun_out/fcomb_ts.log 2>&1; echo "ts rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/run_fcomb.py 64 16 > gpurun_out/fcomb_v4.log 2>&1; echo "v4 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -25 gpurun_out/t_fcomb.log; cat gpurun_out/fcomb_ts.log gpurun_out/fcomb_v4.log
