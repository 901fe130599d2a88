#!/bin/bash
# Round-2 one-GPU call: new persistent cross-attention kernel first (own process), then the GPU suite, timings, ncu.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "cross" > gpurun_out/pytest_cross.log 2>&1; echo "cross exit $?"; tail -6 gpurun_out/pytest_cross.log
timeout 200 python tools/cross_prof.py > gpurun_out/cross_prof.log 2>&1; echo "cross_prof exit $?"; cat gpurun_out/cross_prof.log | tail -8
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-clip --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-3500; tail -5 gpurun_out/bench.err
timeout 120 python tools/attn_prof.py > gpurun_out/attn_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_v8 -s 2 -c 1 -o gpurun_out/attn_v8 python tools/attn_prof.py > gpurun_out/attn_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/attn_ncu.log
