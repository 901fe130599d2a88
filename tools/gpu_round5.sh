#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/attn_tune.py 32760 0,1 > gpurun_out/attn_tune2.log 2>&1; echo "attn_tune exit $?"; cat gpurun_out/attn_tune2.log
timeout 120 python tools/norm_bench.py > gpurun_out/norm_bench.log 2>&1; echo "norm exit $?"; cat gpurun_out/norm_bench.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-clip --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-1800; tail -5 gpurun_out/bench.err
