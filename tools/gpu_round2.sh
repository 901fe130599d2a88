#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, then the ncu launch list of one eager denoise step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-3000; tail -3 gpurun_out/bench.err
SHORT="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-vae"
timeout 300 $SHORT > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 2400 --csv --log-file gpurun_out/step_launches.csv $SHORT > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
python tools/launch_summary.py gpurun_out/step_launches.csv | head -20
