#!/bin/bash
# round profile set: launch list of the bench command's timed step + full captures of the dominant kernels.
# Reports are converted to CSV on the box (gpurun_out/ is capped at 64 MiB); only the small fcomb report travels.
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt gpurun_out/prof_*.ncu-rep gpurun_out/prof_*.raw.csv gpurun_out/launches.csv
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --timed-only"
timeout 300 $B > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "ncu-launches rc=$?" >> gpurun_out/rc.txt
timeout 600 ncu --set full --clock-control none -k regex:conv_tc_kernel\|conv_rs_kernel\|conv_first_tc_kernel -s 32 -c 32 -o gpurun_out/prof_conv $B > gpurun_out/ncu_full.log 2>&1; echo "ncu-conv rc=$?" >> gpurun_out/rc.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fcomb_ts -s 1 -c 1 -o gpurun_out/prof_fcomb $B > gpurun_out/ncu_fcomb.log 2>&1; echo "ncu-fcomb rc=$?" >> gpurun_out/rc.txt
timeout 300 ncu --set full --clock-control none -k regex:gather_\|scatter_\|finalize_\|gauss_head -c 10 -o gpurun_out/prof_gather $B > gpurun_out/ncu_gather.log 2>&1; echo "ncu-gather rc=$?" >> gpurun_out/rc.txt
if [ "$1" == "train" ]; then
timeout 600 ncu --set full --clock-control none -k regex:wgrad_tc_kernel -s 4 -c 6 -o gpurun_out/prof_wgrad python tests/tools/bench_train.py 8 1 --bf16 > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu-wgrad rc=$?" >> gpurun_out/rc.txt
fi
for r in gpurun_out/prof_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep gpurun_out/*.raw.csv
find gpurun_out -name "*.ncu-rep" -size +8M -delete
cat gpurun_out/rc.txt; du -sh gpurun_out
