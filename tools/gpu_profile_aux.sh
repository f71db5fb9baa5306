#!/bin/bash
# Fine-tune GEMM epilogues: CUDA-event timings, then one ncu --set full capture of each gemm2 launch (stage A and B).
mkdir -p gpurun_out
python tools/bench_kernels.py aux > gpurun_out/aux_times.log 2>&1
cat gpurun_out/aux_times.log
BK_ONCE=1 ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16_kernel|gelu_bwd' -c 12 -o gpurun_out/prof_aux -f python tools/bench_kernels.py aux > gpurun_out/ncu_aux.log 2>&1
tail -3 gpurun_out/ncu_aux.log
