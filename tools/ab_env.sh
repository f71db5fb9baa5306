#!/bin/bash
# Same-box A/B of an environment switch:  tools/ab_env.sh VAR=value [rounds]   (A = with the setting, B = without)
SET=$1; R=${2:-2}
for i in $(seq 1 $R); do
  for which in A B; do
    if [ $which = A ]; then E="env $SET"; else E="env"; fi
    $E timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/abe_$which$i.log 2> gpurun_out/abe_$which$i.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/abe_$which$i.log").read().strip().splitlines()[-1])
k=d["kernels"]
print("$which$i", "$SET" if "$which"=="A" else "default", round(d["value"],2), "steps/s", round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"],2), "|", " ".join(f"{n.split('[')[0][:10]}{n[n.find('['):][:8]}={v['ms_per_step']:.3f}" for n,v in list(k.items())[:8]), d["clocks"]["sm_mhz"])
PY
  done
done
