"""Interleaved, repeated timing of the self-attention kernel against torch SDPA at the benchmark shape (B=3) in random order,
so that clock / power drift of the power-capped part hits both alike (timing one after the other favours whichever runs
first). A development build that exports sa_dev_attn_variant can pass a comma list of kernel variants as argv[1]."""
import random
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from stableavatar_b200 import _lib, ops  # noqa: E402

variants = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 and sys.argv[1] else ["ours"]
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B, L = 3, 32760
torch.manual_seed(0)
q, k, v = (torch.randn(B, L, 12, 128, device="cuda").bfloat16() for _ in range(3))
flops = 4.0 * B * L * L * 12 * 128


def run(var, iters=6):
    if var == "sdpa":
        fn = lambda: F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))  # noqa: E731
    else:
        if var != "ours":
            _lib.lib().sa_dev_attn_variant(var)
        fn = lambda: ops.flash_attn(q, k, v)  # noqa: E731
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {v_: [] for v_ in ["sdpa"] + variants}
for _ in range(2):
    for v_ in res:
        run(v_, 2)
random.seed(1)
order = list(res)
for r in range(rounds):
    random.shuffle(order)
    for v_ in order:
        res[v_].append(run(v_))
for v_, ts in res.items():
    ts = sorted(ts)
    med = ts[len(ts) // 2]
    print(f"{str(v_):>6}: median {med:.3f} ms = {flops / med / 1e9:.0f} TFLOP/s   min {ts[0]:.3f}  max {ts[-1]:.3f}", flush=True)
