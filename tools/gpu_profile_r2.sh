#!/bin/bash
# Round-2 profile capture of the CURRENT build (run under gpurun, ONE GPU): plain run first (must exit 0), then the ncu
# launch list of the same command, then `--set full` of one forward's worth of the tensor-core kernels.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-times --graph off"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2d_launches.csv $CMD > gpurun_out/r2_ncu_launch.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:'mlp_fused_kernel|gemm2_bf16_kernel|gemm_bf16_kernel|window_attention_tc' -s 61 -c 18 -o gpurun_out/r2d_prof -f $CMD > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out | tail -6
