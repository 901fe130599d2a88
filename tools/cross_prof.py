"""Full-size fused cross-attention launch (B=3, L=32760, 12 heads; text 512, CLIP 257, audio 21 x 15) for timing / ncu."""
import sys
import torch
sys.path.insert(0, ".")
from stableavatar_b200 import ops

B, L, H, G, A = 3, 32760, 12, 21, 15
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()  # noqa: E731
q = mk(B, L, H, 128)
sets = [(mk(B, 512, H, 128), mk(B, 512, H, 128), 0), (mk(B, 257, H, 128), mk(B, 257, H, 128), 0),
        (mk(B, G * A, H, 128), mk(B, G * A, H, 128), A)]
out = torch.empty_like(q)
for _ in range(3):
    ops.cross_attn3(q, sets, out=out, rows_per_group=L // G)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.cross_attn3(q, sets, out=out, rows_per_group=L // G)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
flop = 4 * L * (512 + 257 + A) * H * 128 * B
print(f"fused cross-attention: {ms:.3f} ms, {flop / ms / 1e9:.0f} TFLOP/s algorithmic")
for name, (k, v, w) in zip(("text", "clip", "audio"), sets):
    e0.record()
    for _ in range(10):
        ops.cross_attn3(q, [(k, v, w)], out=out, rows_per_group=L // G)
    e1.record()
    torch.cuda.synchronize()
    print(f"  {name} alone: {e0.elapsed_time(e1) / 10:.3f} ms")
e0.record()
for _ in range(10):
    ops.flash_attn(q, sets[0][0], sets[0][1], out=out)
e1.record()
torch.cuda.synchronize()
print(f"  text through the self-attention kernel (v8): {e0.elapsed_time(e1) / 10:.3f} ms")
