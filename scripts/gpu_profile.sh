#!/bin/bash
# round profile set: launch list of the bench command + full captures of the dominant kernels
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu-launches rc=$?" >> gpurun_out/rc.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel\|conv_rs_kernel -s 22 -c 14 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu-conv rc=$?" >> gpurun_out/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fcomb_tc6 -s 1 -c 1 -o gpurun_out/prof_fcomb python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fcomb.log 2>&1; echo "ncu-fcomb rc=$?" >> gpurun_out/rc.txt
timeout 900 ncu --set full --clock-control none -k regex:gather_\|scatter_\|finalize_ -c 8 -o gpurun_out/prof_gather python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1; echo "ncu-gather rc=$?" >> gpurun_out/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 4 -c 6 -o gpurun_out/prof_wgrad python scripts/bench_train.py 8 1 --bf16 > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu-wgrad rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; ls -la gpurun_out/*.ncu-rep
