#!/bin/bash
# One gpurun call: tests, smoke, bench, then ncu launch list of one timed step + full captures of the top kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -2 gpurun_out/bench.log | cut -c1-1800; tail -3 gpurun_out/bench.err
SHORT="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$SHORT > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 2100 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
ls -la gpurun_out | head -30
