#!/bin/bash
# e2e pipeline depth in band mode: tools/gpu_exp_e2e_depth.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
for d in 2 2; do
  timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-kernel-times --no-cpu-baseline --e2e-depth $d > gpurun_out/e2e_depth${d}_$N.log 2> gpurun_out/e2e_depth${d}_$N.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_depth${d}_$N.log").read().strip().splitlines()[-1])
print("bands x$N depth $d:", round(d["value"],2), "steps/s", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"],2), "=", round(d["e2e"]["ms_per_step"],3), "ms", d["band_check"] and d["band_check"]["bit_identical_to_unsharded"], d["clocks"]["sm_mhz"])
PY
done
