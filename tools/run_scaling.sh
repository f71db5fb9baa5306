#!/bin/bash
# Scaling record on one box: bands (default) and replicas at N GPUs.  Usage: tools/run_scaling.sh N
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 $TR tools/run_bands_check.py 2>&1 | grep -E "bands x|rror" | tail -3
echo "check wall $(( $(date +%s) - t0 )) s"
for mode in bands replicas; do
  t0=$(date +%s)
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-kernel-times --mode $mode > gpurun_out/scale_${mode}_$N.log 2> gpurun_out/scale_${mode}_$N.err
  echo "$mode rc $? wall $(( $(date +%s) - t0 )) s"
  grep -v "^\*\*\*\|OMP_NUM" gpurun_out/scale_${mode}_$N.err | tail -3
  python - <<PY
import json
d=json.loads(open("gpurun_out/scale_${mode}_$N.log").read().strip().splitlines()[-1])
print("$mode x$N:", round(d["value"],2), "steps/s", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"],2), d["scaling"], d["config"]["launch"], d["clocks"])
PY
done
