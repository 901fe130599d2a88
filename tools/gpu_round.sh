#!/bin/bash
# One gpurun call: smoke, bench, then ncu launch list + one full capture of the dominant kernel (after the plain run exits 0).
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
SHORT="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$SHORT > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:flash_attn_d128 -s 30 -c 2 -o gpurun_out/prof_attn $SHORT > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 60 -c 4 -o gpurun_out/prof_gemm $SHORT > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
ls -la gpurun_out
