"""Probe (torchrun, >= 2 GPUs): can a rank map a peer's device buffer (CUDA IPC through torch's storage sharing) and
write it over NVLink from a plain kernel, and how do a direct peer copy and NCCL all_to_all_single compare on the
sequence-parallel exchange sizes?   torchrun --nproc-per-node 2 tools/ipc_probe.py"""
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())

n = 128 * 1024 * 1024                                    # 256 MB of bf16
recv = torch.zeros(n, device=dev, dtype=torch.bfloat16)
send = torch.full((n,), float(rank + 1), device=dev, dtype=torch.bfloat16)
meta = recv.untyped_storage()._share_cuda_()
metas = [None] * world
dist.all_gather_object(metas, meta)
peers = {}
for r in range(world):
    if r == rank:
        continue
    st = torch.UntypedStorage._new_shared_cuda(*metas[r])
    peers[r] = torch.empty(0, device=st.device, dtype=torch.bfloat16).set_(st, 0, (n,))   # lives on the peer's GPU
dist.barrier()
torch.cuda.synchronize()
nxt = (rank + 1) % world
print(f"[{rank}] mapped peers {list(peers)}; can_access_peer({nxt}) = {torch.cuda.can_device_access_peer(dev.index, nxt)}", flush=True)

# correctness: write my value into the next rank's buffer
peers[nxt].copy_(send)
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
want = float((rank - 1) % world + 1)
print(f"[{rank}] peer write ok = {bool((recv == want).all())}", flush=True)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for mb in (32, 256):
    m = mb * 1024 * 1024 // 2
    ms = timeit(lambda: peers[nxt][:m].copy_(send[:m]))
    print(f"[{rank}] direct peer copy {mb} MB: {ms:.3f} ms = {mb / 1024 / ms * 1e3:.0f} GiB/s", flush=True)
    per = m // world
    splits = [per] * world
    ms = timeit(lambda: dist.all_to_all_single(recv[:per * world], send[:per * world], splits, splits))
    sent = per * 2 * (world - 1) / 2**20
    print(f"[{rank}] nccl all_to_all_single {mb} MB total ({sent:.0f} MB leave the GPU): {ms:.3f} ms = {sent / 1024 / ms * 1e3:.0f} GiB/s out", flush=True)
dist.barrier()
dist.destroy_process_group()
