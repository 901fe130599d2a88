#!/bin/bash
# One gpurun call: tests, smoke, bench, then ncu launch list of one eager step + one full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-2500; tail -3 gpurun_out/bench.err
SHORT="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-vae"
timeout 300 $SHORT > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_v -s 8 -c 1 -o gpurun_out/prof_selfattn_bench $SHORT > gpurun_out/ncu_attn_bench.log 2>&1
echo "ncu attn exit $?"
ls -la gpurun_out | head -30
