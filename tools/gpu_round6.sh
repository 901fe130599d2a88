#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --maxfail=5 > gpurun_out/pytest_k.log 2>&1; echo "kernels exit $?"; tail -5 gpurun_out/pytest_k.log
timeout 300 python tools/attn_time.py > gpurun_out/attn_time.log 2>&1; echo "attn_time exit $?"; cat gpurun_out/attn_time.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-clip --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-1800; tail -5 gpurun_out/bench.err
