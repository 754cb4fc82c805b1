#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_train_host.py 16 > gpurun_out/prof_train_host16.log 2>&1; echo "rc=$?"
grep -v Warn gpurun_out/prof_train_host16.log | head -32
