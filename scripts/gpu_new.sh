#!/bin/bash
# new rows of this session: dataset / views / latent-grid tests, then the whole suite, then the config-5 sweep
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_dataset.py -m gpu -q --no-header -rf -x > gpurun_out/t_dataset.log 2>&1; echo "dataset rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --config 5 --size 512 > gpurun_out/sweep512_n1.log 2>&1; echo "sweep rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_dataset.log; tail -4 gpurun_out/t_all.log; grep -v Warn gpurun_out/sweep512_n1.log | tail -12
