#!/bin/bash
# Round-2 one-GPU call: GPU tests, smoke, library baseline, bench (short), then the ncu launch list of one eager step.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 600 python tools/lib_baseline.py --steps 3 --warmup 2 --out gpurun_out/lib_baseline.json > gpurun_out/lib_baseline.log 2>&1; echo "lib exit $?"; tail -2 gpurun_out/lib_baseline.log | cut -c1-1500
timeout 300 python tools/attn_tune.py > gpurun_out/attn_tune.log 2>&1; echo "attn_tune exit $?"; cat gpurun_out/attn_tune.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-6000; tail -5 gpurun_out/bench.err
