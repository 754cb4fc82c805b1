#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt gpurun_out/prof_*.ncu-rep
timeout 300 python scripts/run_fcomb.py 16 16 > gpurun_out/fcomb_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fcomb_tc5 -s 2 -c 1 -o gpurun_out/prof_fcomb python scripts/run_fcomb.py 16 16 > gpurun_out/ncu_fcomb.log 2>&1; echo "ncu rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt gpurun_out/fcomb_plain.log
