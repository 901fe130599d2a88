"""Audio-adapter cross-attention (sa_attn_small_q) at the benchmark shape: 63 (3 CFG samples x 21 latent frames) x 8 heads of
192, 15 audio queries x 1560 video keys; accuracy vs fp32 torch SDPA, time vs the one-pass HBM bound."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from stableavatar_b200 import ops  # noqa: E402
torch.manual_seed(0)
for (Bf, Lq, Lk, H, D) in [(63, 15, 1560, 8, 192), (63, 17, 1557, 8, 192), (9, 15, 3600, 8, 640)]:
    q = torch.randn(Bf, Lq, H, D, device="cuda").bfloat16()
    k = torch.randn(Bf, Lk, H, D, device="cuda").bfloat16()
    v = torch.randn(Bf, Lk, H, D, device="cuda").bfloat16()
    ref = F.scaled_dot_product_attention(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float()).transpose(1, 2)
    o = ops.attn_small_q(q, k, v)
    err = ((o.float() - ref).norm() / ref.norm()).item()
    for _ in range(3):
        ops.attn_small_q(q, k, v)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attn_small_q(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (k.numel() + v.numel()) * 2 / 1e9
    print(f"B*frames={Bf} q={Lq} keys={Lk} heads={H}x{D}: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s of K+V ({gb * 1e3:.0f} MB), rel-L2 vs fp32 SDPA {err:.2e}")
