#!/bin/bash
# Fine-tune data-parallel record at N GPUs (BASELINE configs[4]):  tools/run_ft_scaling.sh N [variants...]
N=${1:-2}; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
[ "$N" = 1 ] && TR="python"
run() {  # tag, env..., -- bench args
  tag=$1; shift
  t0=$(date +%s)
  env "$@" timeout 400 $TR bench.py --gpus $N --mode finetune --steps 10 --warmup 3 --no-kernel-times $EXTRA > gpurun_out/ft_${tag}_$N.log 2> gpurun_out/ft_${tag}_$N.err
  rc=$?
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ft_${tag}_$N.log").read().strip().splitlines()[-1])
    print("finetune x$N [$tag]:", round(d["ms_per_step"],2), "ms/step", round(d["value"],3), "steps/s; e2e", round(d["e2e"]["ms_per_step"],2), "ms;", d["config"]["parallelism"], d["clocks"])
except Exception as e:
    print("finetune x$N [$tag]: FAILED rc=$rc", e)
PY
  echo "  wall $(( $(date +%s) - t0 )) s"
}
for v in "${@:-flat}"; do
  case $v in
    flat)    EXTRA="--bucket-mb 2048" run flat A=1 ;;
    b64)     EXTRA="--bucket-mb 64" run b64 A=1 ;;
    b64sms)  EXTRA="--bucket-mb 64" run b64sms PANGU_B200_SMS=132 NCCL_MAX_NCHANNELS=8 ;;
    buckets) EXTRA="--dp buckets" run buckets A=1 ;;
  esac
done
