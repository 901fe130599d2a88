"""A few launches of the halo conv kernel at one decoder shape, for ncu."""
import sys
import torch
sys.path.insert(0, ".")
from stableavatar_b200 import ops
cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (96, 96)
T, H, W = (4, 480, 832) if cout == 96 else (4, 240, 416)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(T + 2, H, W, cin, device="cuda", generator=g).bfloat16()
w5 = (torch.randn(cout, 3, 3, 3, cin, device="cuda", generator=g) * (27 * cin) ** -0.5).bfloat16()
bias = torch.randn(cout, device="cuda", generator=g)
wp = ops.pack_conv_weight_halo(w5)
out = torch.empty(T, H, W, cout, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.conv3d_halo_cl(x, wp, bias, cout=cout, out=out)
torch.cuda.synchronize()
print("ok")
