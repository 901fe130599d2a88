"""Full-size check + timing of the NVLink peer exchange (csrc/sp_exchange.cu) against the NCCL all_to_all_single path on
the same data (torchrun, P GPUs): the receive buffers hold exactly what NCCL delivers (Q / KV / O), the fused
norm+RoPE+scatter equals norm-then-scatter, the per-sample pipelined attention region equals the serial one bit for bit;
then ms per exchange for NCCL / peer, and the attention region serial vs pipelined (the exposed exchange time).
  torchrun --nproc-per-node 2 tools/sp_peer_check.py [L] [heads]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from stableavatar_b200 import ops, sequence_parallel as sp  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
ops.sp_set_barrier_timeout_ms(60_000)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32760
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 12
B, d = 3, 128
Ll = (L + world - 1) // world
pl = sp.plan(nh, world, rank)
n_src = len(pl.q_sources)
g = torch.Generator(device=dev).manual_seed(rank)
qkv = torch.randn(B * Ll, 3 * nh * d, device=dev, generator=g).bfloat16()
q5 = qkv.view(B, Ll, 3, nh, d)
C = nh * d
ok_all = True


def log(*a):
    print(f"[{rank}]", *a, flush=True)


def check(name, flag):
    global ok_all
    ok_all &= bool(flag)
    log(f"{name}: {'equal' if flag else 'DIFFERENT'}")


# NCCL path: Q [Lq, B, hp, d], KV [L, B, 2, hp, d] (token-major); peer path: batch-outermost
Qn, KVn = sp.exchange_qkv(pl, q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], None, kv=q5[:, :, 1:3])
torch.cuda.synchronize()
px = sp.PeerExchange(pl, B, Ll, nh, d, dev)
px.scatter_qkv(qkv, None)
px.barrier(0)
torch.cuda.synchronize()
check("peer qkv scatter vs NCCL: Q", torch.equal(px.q_recv, Qn.transpose(0, 1)))
check("peer qkv scatter vs NCCL: KV", torch.equal(px.kv_recv, KVn.transpose(0, 1)))
# producer fusion: RMSNorm + RoPE inside the scatter == sa_rmsnorm_rope in place followed by the plain scatter
wq = (1 + 0.1 * torch.randn(C, device=dev, generator=g)).bfloat16()
wk = (1 + 0.1 * torch.randn(C, device=dev, generator=g)).bfloat16()
grid = (21, 30, 52)
ang = torch.rand(1024, 64, device=dev, generator=g) * 6.28
freqs = torch.stack([ang.cos(), ang.sin()], -1).float().contiguous()
norm = (wq, wk, freqs, grid, rank * Ll)
ref = qkv.clone()
ops.rmsnorm_rope_(ref[:, :C], wq, ref[:, C:2 * C], wk, freqs=freqs, grid=grid, rows_per_batch=Ll, tok_offset=rank * Ll)
px.scatter_qkv(ref, None)
px.barrier(0)
torch.cuda.synchronize()
Qr, KVr = px.q_recv.clone(), px.kv_recv.clone()
dist.barrier()
for b in range(B):                                   # sample by sample, as the pipeline issues it
    px.scatter_qkv(qkv, norm, b, 1)
px.barrier(0)
torch.cuda.synchronize()
check("fused norm+rope+scatter (per sample) vs norm then scatter: Q", torch.equal(px.q_recv, Qr))
check("fused norm+rope+scatter (per sample) vs norm then scatter: KV", torch.equal(px.kv_recv, KVr))
dist.barrier()
O = torch.randn(n_src * Ll, B, pl.hp, d, device=dev, generator=g).bfloat16()
On = sp.exchange_out(pl, O, B, Ll, nh, d, None).contiguous()
px.o_send.copy_(O.transpose(0, 1))
px.scatter_o()
px.barrier(0)
torch.cuda.synchronize()
check("peer O scatter vs NCCL", torch.equal(px.o_recv, On))
dist.barrier()
a = px.attention(qkv, norm, pipelined=False, fused_o=False).clone()
torch.cuda.synchronize()
dist.barrier()
b_ = px.attention(qkv, norm, pipelined=True, fused_o=False).clone()
torch.cuda.synchronize()
check("attention region pipelined vs serial", torch.equal(a, b_))
dist.barrier()
c_ = px.attention(qkv, norm, pipelined=False, fused_o=True).clone()
torch.cuda.synchronize()
check("attention region with O stored to the owners by the attention epilogue (TMA) vs scatter kernel", torch.equal(a, c_))
dist.barrier()
d_ = px.attention(qkv, norm, pipelined=True, fused_o=True).clone()
torch.cuda.synchronize()
check("same, pipelined per sample", torch.equal(a, d_))


def timeit(fn, iters=10):
    for _ in range(2):
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


def peer_qkv(nm):
    px.scatter_qkv(qkv, nm)
    px.barrier(0)


def peer_o():
    px.scatter_o()
    px.barrier(0)


out_mb = (qkv.numel() * 2 / 3 * (2 * pl.qs + 1)) * (world - 1) / world / 2**20      # bytes leaving this GPU, roughly
t_n = timeit(lambda: sp.exchange_qkv(pl, q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], None, kv=q5[:, :, 1:3]))
t_p = timeit(lambda: peer_qkv(None))
log(f"qkv exchange: nccl (pack + 2 all_to_all) {t_n:.3f} ms, peer (scatter + barrier) {t_p:.3f} ms, ~{out_mb:.0f} MB leave the GPU -> "
    f"{out_mb / 1024 / t_p * 1e3:.0f} GiB/s")
t_f = timeit(lambda: peer_qkv(norm))
log(f"fused norm+rope+scatter + barrier {t_f:.3f} ms")
t_n = timeit(lambda: sp.exchange_out(pl, O, B, Ll, nh, d, None).contiguous())
t_p = timeit(peer_o)
log(f"O exchange: nccl (all_to_all + unpack) {t_n:.3f} ms, peer {t_p:.3f} ms")
t_attn = timeit(lambda: ops.flash_attn(px.q_recv, px.kv_recv[:, :, 0], px.kv_recv[:, :, 1], out=px.o_send))
t_s0 = timeit(lambda: px.attention(qkv, norm, pipelined=False, fused_o=False))
t_s = timeit(lambda: px.attention(qkv, norm, pipelined=False, fused_o=True))
t_pp = timeit(lambda: px.attention(qkv, norm, pipelined=True, fused_o=True))
log(f"attention alone {t_attn:.3f} ms; region serial with scatter_o kernel {t_s0:.3f} ms (exposed {t_s0 - t_attn:.3f}); serial with "
    f"fused O store {t_s:.3f} ms (exposed {t_s - t_attn:.3f}); pipelined per CFG sample + fused O {t_pp:.3f} ms (exposed {t_pp - t_attn:.3f})")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
