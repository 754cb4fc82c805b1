#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do PMU_CONV_DEBUG=$d timeout 200 python scripts/time_convs_small.py > gpurun_out/convs_small_dbg$d.log 2>&1; cat gpurun_out/convs_small_dbg$d.log; done
