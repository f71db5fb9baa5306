#!/bin/bash
# K/V-only halos vs whole-qkv halos in band mode: tools/gpu_exp_halo_kv.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bands_nccl.py -x -q 2>&1 | tail -2
for kv in 0 1 0 1; do
  PANGU_B200_HALO_KV=$kv timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-kernel-times --no-cpu-baseline > gpurun_out/halo_kv${kv}_$N.log 2> gpurun_out/halo_kv${kv}_$N.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/halo_kv${kv}_$N.log").read().strip().splitlines()[-1])
print("bands x$N halo_kv=$kv:", round(d["value"],2), "steps/s", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"],2), d["band_check"] and d["band_check"]["bit_identical_to_unsharded"], d["clocks"]["sm_mhz"])
PY
done
