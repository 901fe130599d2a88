"""HBM roofline of the row kernels at the benchmark's size (B*L = 98 280 rows x 1536): LayerNorm + AdaLN modulate,
LayerNorm affine, RMSNorm + RoPE (q and k of a fused QKV buffer). Algorithmic bytes = one read + one write of the rows."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, ".")
from stableavatar_b200 import ops  # noqa: E402

peak = 6552.0
pk = Path("MEASURED_PEAKS.json")
if pk.exists():
    peak = json.loads(pk.read_text())["hbm_gbs"]
B, L, C = 3, 32760, 1536
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B * L, C, device="cuda", generator=g).bfloat16()
e = (torch.randn(B, 6 * C, device="cuda", generator=g) * 0.1).bfloat16()
w = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).bfloat16()
b = (0.1 * torch.randn(C, device="cuda", generator=g)).bfloat16()
qkv = torch.randn(B * L, 3 * C, device="cuda", generator=g).bfloat16()
ang = torch.rand(1024, 64, device="cuda", generator=g) * 6.28
freqs = torch.stack([ang.cos(), ang.sin()], -1).float().contiguous()
out = torch.empty_like(x)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                                   # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


rows = x.numel() * 2
cases = [("layernorm + modulate (norm1 / norm2)", lambda: ops.layernorm(x, shift=e[:, :C], scale=e[:, C:2 * C], mod_bs=6 * C, rows_per_batch=L, out=out), 2 * rows),
         ("layernorm affine (norm3)", lambda: ops.layernorm(x, weight=w, bias=b, out=out), 2 * rows),
         ("rmsnorm + rope (q, k in place)", lambda: ops.rmsnorm_rope_(qkv[:, :C], w, qkv[:, C:2 * C], w, freqs=freqs, grid=(21, 30, 52), rows_per_batch=L), 4 * rows)]
for name, fn, nbytes in cases:
    ms = timeit(fn)
    gbs = nbytes / ms / 1e6
    print(f"{name}: {ms:.4f} ms, {gbs:.0f} GB/s = {gbs / peak:.2f} of the measured {peak:.0f} GB/s HBM copy", flush=True)
