#!/bin/bash
# Window-attention backward: CUDA-event timings (both stages, un-rolled / rolled), then one ncu --set full capture per stage.
mkdir -p gpurun_out
python tools/bench_kernels.py attnbwd > gpurun_out/attnbwd_times.log 2>&1
cat gpurun_out/attnbwd_times.log
BK_ONCE=1 ncu --set full --clock-control none --import-source on -k regex:'window_attention_bwd' -c 4 -o gpurun_out/prof_attnbwd -f python tools/bench_kernels.py attnbwd > gpurun_out/ncu_attnbwd.log 2>&1
tail -2 gpurun_out/ncu_attnbwd.log
