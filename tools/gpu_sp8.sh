#!/bin/bash
# 8-GPU call (gpurun --gpus 8): SP parity tests at world 4 and 8 (world 2 runs in tools/gpu_sp.sh), pipelined VAE, bench N=8.
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
T=tests/test_sp_gpu.py::test_sp_forward_equals_single_gpu
H=tests/test_sp_gpu.py::test_sp_forward_40_heads_equals_single_gpu
timeout 1200 python -m pytest -m gpu -q "$T[8-peer]" "$T[8-peer_pipelined]" "$T[8-peer_serial_unfused]" "$T[8-nccl]" "$T[4-peer]" "$T[4-nccl]" \
  "$H[8]" "$H[4]" "tests/test_vae_gpu.py::test_pipeline_parallel_decode_equals_single_gpu[4]" > gpurun_out/pytest_sp_n8.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_sp_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench exit $?"
tail -1 gpurun_out/bench_n8.log | cut -c1-6000; grep -v "^\[W\|NCCL\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_n8.err | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/sp_peer_check.py > gpurun_out/sp_peer_check_n8.log 2>&1; echo "peer check exit $?"; grep "^\[0\]" gpurun_out/sp_peer_check_n8.log | tail -12
