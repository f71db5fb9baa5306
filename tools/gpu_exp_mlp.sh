#!/bin/bash
# fused-Mlp kernel alone, L2 flushed between launches: tools/gpu_exp_mlp.sh "<lib tags>"   (cur = in-tree build)
for v in ${1:-cur}; do
  if [ $v = cur ]; then unset PANGU_B200_LIB; else export PANGU_B200_LIB=$PWD/ab/libpangu_$v.so; fi
  for C in 192 384; do
    timeout 120 python - <<PY
import os, sys, torch
sys.path.insert(0, "pangu-pytorch-demo_b200")
from pangu_b200 import ops
C=$C; M = 521280 if C == 192 else 131040
g = torch.Generator(device="cuda").manual_seed(0)
xb = torch.randn(M, C, device="cuda", generator=g).bfloat16(); x = torch.randn(M, C, device="cuda", generator=g)
w1 = (torch.randn(4*C, C, device="cuda", generator=g)*0.05).bfloat16(); w2 = (torch.randn(C, 4*C, device="cuda", generator=g)*0.05).half()
b1, b2 = torch.zeros(4*C, device="cuda"), torch.zeros(C, device="cuda"); ga, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
big = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
def run(): ops.mlp_ln_residual_bf16(xb, w1, b1, w2, b2, ga, be, x)
for _ in range(3): run()
ts=[]
for _ in range(10):
    big.zero_()                                   # flush L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print("lib=$v C=%d: median %.4f ms  min %.4f" % (C, ts[len(ts)//2], ts[0]))
PY
  done
done 2>&1 | grep lib= | tee gpurun_out/exp_mlp.log
