#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -1 gpurun_out/bench.log | cut -c1-1200; tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 3200 --csv --log-file gpurun_out/step_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-vae --no-clip --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_halo -s 2 -c 1 -o gpurun_out/conv_halo96_v2 -f python tools/conv_prof.py 96 96 > gpurun_out/conv_ncu2.log 2>&1; echo "ncu conv exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vae_launches3.csv python tools/vae_bench.py 21 > gpurun_out/vae_ncu.log 2>&1; tail -1 gpurun_out/vae_ncu.log
