#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do
  PMU_CONV_DEBUG=$d timeout 300 python scripts/time_convs.py 64 > gpurun_out/time_convs_dbg$d.log 2>&1
done
paste -d'|' <(cut -c1-72 gpurun_out/time_convs_dbg0.log) <(cut -c41-72 gpurun_out/time_convs_dbg1.log) <(cut -c41-72 gpurun_out/time_convs_dbg2.log) <(cut -c41-72 gpurun_out/time_convs_dbg3.log)
