#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt gpurun_out/prof_fcomb6.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fcomb_tc6 -s 2 -c 1 -o gpurun_out/prof_fcomb6 python scripts/run_fcomb.py 16 16 > gpurun_out/ncu_fcomb6.log 2>&1; echo "ncu rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -3 gpurun_out/ncu_fcomb6.log
