#!/bin/bash
# What the driver runs at round end, on one box: smoke, the GPU tests, the default bench line, the reference arm.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout 600 python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/final_bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/final_bench_ref.log
