#!/bin/bash
# Round-1 (final) profile capture of the CURRENT build (run under gpurun): plain bench, ncu launch list, --set full of the
# top kernels (fused MLP, pair GEMM, GEMM+LN, tcgen05 window attention).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-times --graph off"
$CMD > gpurun_out/plain_e.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_e.csv $CMD > gpurun_out/ncu_launch_e.log 2>&1
$CMD > gpurun_out/plain_e2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'mlp_fused_kernel|gemm2_bf16_kernel|gemm_bf16_kernel|window_attention_tc' -s 71 -c 24 -o gpurun_out/prof_r1d -f $CMD > gpurun_out/ncu_full_d.log 2>&1
ls -la gpurun_out | tail -8
