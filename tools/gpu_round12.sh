#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-900; tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 3200 --csv --log-file gpurun_out/step_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-vae --no-clip --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
