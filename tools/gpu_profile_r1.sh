#!/bin/bash
# Round-1 profile capture (run under gpurun): plain bench, launch list, one --set full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-times"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_kernel|window_attention_bf16' -s 87 -c 30 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
