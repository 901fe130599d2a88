#!/bin/bash
# Multi-GPU call (gpurun --gpus N): peer-exchange check, SP / pipelined-VAE parity tests, bench at N GPUs.
N=${1:-2}
STEPS=${2:-3}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sp_peer_check.py > gpurun_out/sp_peer_check_n$N.log 2>&1; echo "peer check exit $?"; grep -v "^\[W\|NCCL" gpurun_out/sp_peer_check_n$N.log | tail -24
timeout 900 python -m pytest tests/test_sp_gpu.py tests/test_vae_gpu.py -m gpu -q -k "sp_forward or pipeline_parallel" > gpurun_out/pytest_sp_n$N.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_sp_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
tail -1 gpurun_out/bench_n$N.log | cut -c1-6000; grep -v "^\[W\|NCCL\|^$" gpurun_out/bench_n$N.err | tail -8
