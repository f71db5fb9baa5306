#!/bin/bash
# End-of-round record of the CURRENT build (run under gpurun): GPU tests, default bench, fine-tune bench, ncu launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench13.log 2> gpurun_out/bench13.err; echo "bench rc $?"
python bench.py --mode finetune --steps 5 --warmup 3 > gpurun_out/ft3.log 2> gpurun_out/ft3.err; echo "finetune rc $?"; tail -c 900 gpurun_out/ft3.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-times --graph off"
$CMD > gpurun_out/plain_d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_d.csv $CMD > gpurun_out/ncu_launch_d.log 2>&1
tail -2 gpurun_out/ncu_launch_d.log
python tools/bench_kernels.py attnbwd
