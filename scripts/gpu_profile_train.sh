#!/bin/bash
# ncu --set full of the training step's new / dominant kernels (one eager step between cudaProfilerStart / Stop):
# the convolution with the fused BatchNorm statistics, the weight gradient, the BatchNorm backward passes.
mkdir -p gpurun_out; rm -f gpurun_out/prof_train*.ncu-rep gpurun_out/prof_train*.raw.csv
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:conv_rs_kernel\|conv_tc_kernel -c 24 -o gpurun_out/prof_train_conv python scripts/train_launches.py > gpurun_out/ncu_train_conv.log 2>&1; echo "conv rc=$?"
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:wgrad_tc_kernel -c 12 -o gpurun_out/prof_train_wgrad python scripts/train_launches.py > gpurun_out/ncu_train_wgrad.log 2>&1; echo "wgrad rc=$?"
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:bn_bwd_\|bn_act_\|bn_finalize\|fcomb_last\|pack_conv\|unpack_wgrad\|wgrad_smallcin -c 30 -o gpurun_out/prof_train_bn python scripts/train_launches.py > gpurun_out/ncu_train_bn.log 2>&1; echo "bn rc=$?"
for r in gpurun_out/prof_train*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
find gpurun_out -name "prof_train*.ncu-rep" -delete
ls -la gpurun_out/prof_train*.raw.csv; du -sh gpurun_out
