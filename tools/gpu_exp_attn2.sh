#!/bin/bash
# attention variants timed alone (tools/attn_trace.py): tools/gpu_exp_attn2.sh "<tags>"
for v in ${1:-base4 cur base4 cur}; do
  if [ $v = cur ]; then unset PANGU_B200_LIB; else export PANGU_B200_LIB=$PWD/ab/libpangu_$v.so; fi
  for s in A B; do
    timeout 120 python tools/attn_trace.py $s > /tmp/at.log 2>&1
    if grep -q 'ms / launch' /tmp/at.log; then echo "== $v $s: $(grep 'ms / launch' /tmp/at.log | tr '\n' ' ')"; else echo "== $v $s FAILED: $(tail -3 /tmp/at.log | tr '\n' ' ')"; fi
  done
done 2>&1 | tee gpurun_out/exp_attn_variants2.log
