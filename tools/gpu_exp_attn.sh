#!/bin/bash
# Round 2 experiment: attention kernel variants on ONE box (ab/libpangu_<tag>.so vs the in-tree build), after the GPU tests.
#   tools/gpu_exp_attn.sh "<tags>"      e.g. "base cur base cur"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/exp_pytest.log
for v in ${1:-base cur base cur}; do
  if [ $v = cur ]; then unset PANGU_B200_LIB; else export PANGU_B200_LIB=$PWD/ab/libpangu_$v.so; fi
  for s in A B; do
    echo "== $v $s: $(timeout 120 python tools/attn_trace.py $s 2>&1 | grep 'ms / launch' | tr '\n' ' ')"
  done
done 2>&1 | tee gpurun_out/exp_attn_variants.log
unset PANGU_B200_LIB
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp_bench.log 2> gpurun_out/exp_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/exp_bench.log").read().strip().splitlines()[-1])
k=d["kernels"]
print(round(d["value"],2), "steps/s", round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"],2), "|", " ".join(f"{n.split('[')[0][:10]}{n[n.find('['):][:8]}={v['ms_per_step']:.3f}" for n,v in list(k.items())[:8]), d["clocks"]["sm_mhz"])
PY
