"""Full-size check + timing of the NVLink peer exchange (csrc/sp_exchange.cu) against the NCCL all_to_all_single path on
the same data (torchrun, P GPUs): exact equality of Q / KV / O layouts, then ms per exchange for both.
  torchrun --nproc-per-node 2 tools/sp_peer_check.py [L] [heads]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from stableavatar_b200 import sequence_parallel as sp  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32760
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 12
B, d = 3, 128
Ll = (L + world - 1) // world
pl = sp.plan(nh, world, rank)
g = torch.Generator(device=dev).manual_seed(rank)
qkv = torch.randn(B * Ll, 3 * nh * d, device=dev, generator=g).bfloat16()
q5 = qkv.view(B, Ll, 3, nh, d)


def log(*a):
    print(f"[{rank}]", *a, flush=True)


Qn, KVn = sp.exchange_qkv(pl, q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], None, kv=q5[:, :, 1:3])
torch.cuda.synchronize()
log("nccl exchange done", tuple(Qn.shape), tuple(KVn.shape))
px = sp.PeerExchange(pl, B, Ll, nh, d, dev)
torch.cuda.synchronize()
log("peer buffers mapped")
Qp, KVp = px.exchange_qkv(qkv)
torch.cuda.synchronize()
log("peer qkv exchange: Q equal", torch.equal(Qp, Qn), "KV equal", torch.equal(KVp, KVn))
# producer fusion: RMSNorm + RoPE inside the scatter == sa_rmsnorm_rope in place followed by the plain scatter
from stableavatar_b200 import ops  # noqa: E402
C = nh * d
wq = (1 + 0.1 * torch.randn(C, device=dev, generator=g)).bfloat16()
wk = (1 + 0.1 * torch.randn(C, device=dev, generator=g)).bfloat16()
Fg, Hg, Wg = 21, 30, 52
ang = torch.rand(1024, 64, device=dev, generator=g) * 6.28
freqs = torch.stack([ang.cos(), ang.sin()], -1).float().contiguous()
ref = qkv.clone()
ops.rmsnorm_rope_(ref[:, :C], wq, ref[:, C:2 * C], wk, freqs=freqs, grid=(Fg, Hg, Wg), rows_per_batch=Ll, tok_offset=rank * Ll)
Qr, KVr = px.exchange_qkv(ref)
Qr, KVr = Qr.clone(), KVr.clone()
torch.cuda.synchronize()
dist.barrier()
Qf, KVf = px.norm_rope_exchange_qkv(qkv, wq, wk, freqs, (Fg, Hg, Wg), rank * Ll)
torch.cuda.synchronize()
log("fused norm+rope+scatter: Q equal", torch.equal(Qf, Qr), "KV equal", torch.equal(KVf, KVr))
dist.barrier()
O = torch.randn(Qn.shape, device=dev, generator=g).bfloat16()
On = sp.exchange_out(pl, O, B, Ll, nh, d, None).contiguous()
torch.cuda.synchronize()
Op = px.exchange_out(O)
torch.cuda.synchronize()
log("peer O exchange equal", torch.equal(Op, On))


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


out_mb = (qkv.numel() * 2 / 3 * (2 * pl.qs + 1)) * (world - 1) / world / 2**20      # bytes leaving this GPU, roughly
t_n = timeit(lambda: sp.exchange_qkv(pl, q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], None, kv=q5[:, :, 1:3]))
t_p = timeit(lambda: px.exchange_qkv(qkv))
log(f"qkv exchange: nccl (pack + 2 all_to_all) {t_n:.3f} ms, peer (scatter + barrier) {t_p:.3f} ms, ~{out_mb:.0f} MB leave the GPU -> "
    f"{out_mb / 1024 / t_p * 1e3:.0f} GiB/s")
def unfused():
    ops.rmsnorm_rope_(ref[:, :C], wq, ref[:, C:2 * C], wk, freqs=freqs, grid=(Fg, Hg, Wg), rows_per_batch=Ll, tok_offset=rank * Ll)
    px.exchange_qkv(ref)


t_u = timeit(unfused)
t_f = timeit(lambda: px.norm_rope_exchange_qkv(qkv, wq, wk, freqs, (Fg, Hg, Wg), rank * Ll))
log(f"rmsnorm+rope then scatter+barrier {t_u:.3f} ms, fused {t_f:.3f} ms")
t_n = timeit(lambda: sp.exchange_out(pl, O, B, Ll, nh, d, None).contiguous())
t_p = timeit(lambda: px.exchange_out(O))
log(f"O exchange: nccl (all_to_all + unpack) {t_n:.3f} ms, peer {t_p:.3f} ms")
dist.barrier()
dist.destroy_process_group()
