#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 3; do PMU_CONV_DEBUG=$d timeout 200 python scripts/time_convs.py > gpurun_out/time_convs_dbg$d.log 2>&1; echo "debug $d"; grep -E "\.T |TOTAL" gpurun_out/time_convs_dbg$d.log; done
