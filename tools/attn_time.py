"""Standalone timing + accuracy of the self-attention kernel against torch SDPA (the library kernel the reference
dispatches to) on the benchmark's shape: B=1 and B=3, 12 heads, L=32760, d=128."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from stableavatar_b200 import _lib, ops  # noqa: E402

torch.manual_seed(0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32760
# development builds export sa_dev_attn_variant (A/B of kernel variants in one process); the shipped library has one kernel
VARIANTS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [None]


def timeit(fn, iters=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for B in (1, 3):
    q, k, v = (torch.randn(B, L, 12, 128, device="cuda").bfloat16() for _ in range(3))
    flops = 4.0 * B * L * L * 12 * 128
    ref = F.scaled_dot_product_attention(q[:, :2048].transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float()).transpose(1, 2)
    t = timeit(lambda: F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)))
    print(f"B={B} torch SDPA: {t:.3f} ms = {flops / t / 1e9:.0f} TFLOP/s", flush=True)
    base = None
    for var in VARIANTS:
        if var is not None:
            _lib.lib().sa_dev_attn_variant(var)
        o = ops.flash_attn(q, k, v)
        torch.cuda.synchronize()
        err = ((o[:, :2048].float() - ref).norm() / ref.norm()).item()
        same = "" if base is None else f", bit-equal to first variant: {torch.equal(o, base)}"
        base = o if base is None else base
        t = timeit(lambda: ops.flash_attn(q, k, v))
        print(f"B={B} flash_attn variant {var}: {t:.3f} ms = {flops / t / 1e9:.0f} TFLOP/s, rel-L2 vs fp32 SDPA (2048 rows) {err:.2e}{same}", flush=True)
