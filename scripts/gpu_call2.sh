#!/bin/bash
# parity suite, bench, ncu launch list + full capture of the dominant kernel
mkdir -p gpurun_out
rm -f gpurun_out/rc.txt
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu-launches rc=$?" >> gpurun_out/rc.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 31 -c 10 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu-full rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -3 gpurun_out/t_all.log
tail -1 gpurun_out/bench.log
ls -la gpurun_out
