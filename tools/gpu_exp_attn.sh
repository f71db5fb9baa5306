#!/bin/bash
# Round 2 experiment: softmax variants of the tcgen05 attention kernel (ab/libpangu_{base,p0,p3,p4}.so vs the in-tree build)
# and the L2 evict_last hint of the fused block tail, all on ONE box.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/exp_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/exp_pytest.log
for v in base p0 cur p3 p4 base cur; do
  if [ $v = cur ]; then unset PANGU_B200_LIB; else export PANGU_B200_LIB=$PWD/ab/libpangu_$v.so; fi
  for s in A B; do
    echo "== $v $s: $(timeout 120 python tools/attn_trace.py $s 2>&1 | grep 'ms / launch' | tr '\n' ' ')"
  done
done 2>&1 | tee gpurun_out/exp_attn_variants.log
unset PANGU_B200_LIB
tools/ab_env.sh PANGU_MLP_DBG=128 2 2>&1 | tee gpurun_out/exp_l2hint.log
