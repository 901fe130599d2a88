"""Single self-attention-shaped launch for ncu (B=1, H=12, L=32760, d=128)."""
import sys
import torch
sys.path.insert(0, ".")
from stableavatar_b200 import ops
torch.manual_seed(0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32760
q, k, v = (torch.randn(1, L, 12, 128, device="cuda").bfloat16() for _ in range(3))
for _ in range(3):
    o = ops.flash_attn(q, k, v)
torch.cuda.synchronize()
print("ok", o.float().abs().mean().item())
