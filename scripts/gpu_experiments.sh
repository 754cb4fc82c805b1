#!/bin/bash
# One-GPU A/B run of the opt-in kernel variants that were written without GPU time:
#   1. their parity tests (PMU_TEST_EXPERIMENTAL=1 un-skips them),
#   2. the resident step (bench.py --timed-only) with each switch on its own and with all of them.
# Usage: gpurun --timeout 1800 -- 'bash scripts/gpu_experiments.sh'
mkdir -p gpurun_out; rm -f gpurun_out/exp_*.log gpurun_out/exp_rc.txt
# 0. the hardware question behind the f16 fcomb variants, isolated: TMEM read port in register bytes or in columns?
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I probabilistic-multiplanar-unet_b200/csrc \
  -o /tmp/tmem_ld_bench scripts/tmem_ld_bench.cu > gpurun_out/exp_tmem_ld.log 2>&1 && \
  timeout 120 /tmp/tmem_ld_bench >> gpurun_out/exp_tmem_ld.log 2>&1; echo "tmem_ld_bench rc=$?" >> gpurun_out/exp_rc.txt
# one pytest process per variant: a trapped kernel poisons the CUDA context of its own process only
: > gpurun_out/exp_tests.log
for k in "fcomb_softmax_accum_bf16 and tshalf" "fcomb_softmax_accum_bf16 and sshalf" "conv_gemm_pool_bf16" "conv_rs_resident_weights" \
         "convt_resident_weights and RESW" "convt_resident_weights and PAIR"; do
  PMU_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --no-header -rf -k "$k" >> gpurun_out/exp_tests.log 2>&1
  echo "tests [$k] rc=$?" >> gpurun_out/exp_rc.txt
done
for k in accumulate_graphed mc_kl; do
  PMU_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q --no-header -rf -k "$k" >> gpurun_out/exp_tests.log 2>&1
  echo "tests [$k] rc=$?" >> gpurun_out/exp_rc.txt
done
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --timed-only"
run() {   # name, env assignments...
  local name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/exp_$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/exp_$name.log | cut -c1-160)" >> gpurun_out/exp_rc.txt
}
run default PMU_NOOP=1
run fcomb_ts PMU_FCOMB_TS=1
run fcomb_ts_f16 PMU_FCOMB_TS=2
run fcomb_ss_f16 PMU_FCOMB_F16=1
run pool_split PMU_POOL_SPLIT=1
run res128 PMU_CONV_RES128=1
run convt_resw PMU_CONVT_RESW=1
run convt_pair PMU_CONVT_PAIR=1
run all PMU_FCOMB_TS=2 PMU_POOL_SPLIT=1 PMU_CONV_RES128=1 PMU_CONVT_PAIR=1
run default_again PMU_NOOP=1
B0="$B"; B="$B --graph"; run graph PMU_NOOP=1; B="$B0"
# per-layer A/B of the conv variants (median of 5 per layer, CUDA events)
timeout 200 python scripts/time_convs.py > gpurun_out/exp_convs_default.log 2>&1; echo "time_convs default rc=$?" >> gpurun_out/exp_rc.txt
PMU_POOL_SPLIT=1 PMU_CONV_RES128=1 PMU_CONVT_PAIR=1 timeout 200 python scripts/time_convs.py > gpurun_out/exp_convs_variants.log 2>&1; echo "time_convs variants rc=$?" >> gpurun_out/exp_rc.txt
paste -d'|' <(cut -c1-62 gpurun_out/exp_convs_default.log) <(cut -c41-62 gpurun_out/exp_convs_variants.log) > gpurun_out/exp_convs_ab.txt
# slice batch: the 16x16 layers (Cout = 1024) run 512 tiles = 3.46 waves of 148 SMs at batch 64 (13 % tail), 6.9 at 128
B="$B --slice-batch 128"; run batch128 PMU_NOOP=1
B="${B/--slice-batch 128/--slice-batch 256}"; run batch256 PMU_NOOP=1
cat gpurun_out/exp_rc.txt; grep -E 'passed|failed|error' gpurun_out/exp_tests.log; cat gpurun_out/exp_tmem_ld.log gpurun_out/exp_convs_ab.txt
