#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-400; tail -3 gpurun_out/bench.err
