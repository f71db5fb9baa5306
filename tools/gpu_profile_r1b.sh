#!/bin/bash
# Round-1 profile capture of the CURRENT build (run under gpurun): plain bench, launch list, --set full of the top kernels.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-times --graph off"
$CMD > gpurun_out/plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_b.csv $CMD > gpurun_out/ncu_launch_b.log 2>&1
$CMD > gpurun_out/plain_b2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'mlp_fused_kernel|gemm2_bf16_kernel|gemm_bf16_kernel|window_attention_bf16' -s 71 -c 24 -o gpurun_out/prof_r1b -f $CMD > gpurun_out/ncu_full_b.log 2>&1
ls -la gpurun_out | tail -8
