#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
PMU_FCOMB_TS=1 timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -x -k fcomb_softmax > gpurun_out/t_fcomb.log 2>&1; echo "fcomb ts rc=$?" >> gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_train.py -m gpu -q --no-header -rf -x > gpurun_out/t_fcomb2.log 2>&1; echo "layers + train rc=$?" >> gpurun_out/rc.txt
PMU_FCOMB_TS=1 timeout 300 python scripts/run_fcomb.py 64 16 > gpurun_out/fcomb_ts.log 2>&1; echo "ts rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/run_fcomb.py 64 16 > gpurun_out/fcomb_v4.log 2>&1; echo "v4 rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/time_convs.py > gpurun_out/time_convs_epi.log 2>&1; echo "convs rc=$?" >> gpurun_out/rc.txt
timeout 300 python scripts/prof_train_host.py > gpurun_out/prof_train_host.log 2>&1; echo "prof rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -3 gpurun_out/t_fcomb.log; tail -3 gpurun_out/t_fcomb2.log; cat gpurun_out/fcomb_ts.log gpurun_out/fcomb_v4.log; cat gpurun_out/time_convs_epi.log; grep -v Warn gpurun_out/prof_train_host.log | head -40
