import sys; sys.path.insert(0, ".")
import torch
from stableavatar_b200 import fp32_mode as F
def rel(a, b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
g = torch.Generator(device="cuda").manual_seed(0)
for (M,N,K) in ((300,136,1000),(256,256,1536),(64,64,64)):
    x = torch.randn(M, K, device="cuda", generator=g) * 3
    w = torch.randn(N, K, device="cuda", generator=g)
    b = torch.randn(N, device="cuda", generator=g)
    want = x.double() @ w.double().t() + b.double()
    got = F.linear(x, w, b)
    torch.backends.cuda.matmul.allow_tf32 = False
    print(M,N,K, "split:", rel(got, want), "torch fp32:", rel(x @ w.t() + b, want), "bf16:", rel(x.bfloat16().float() @ w.bfloat16().float().t() + b, want))
    s = F.split3(x, 0).float().view(M, 6, -1)
    print("  split terms reconstruct x:", rel(s[:,0]+s[:,2]+s[:,5], x))
    print("  gelu:", rel(F.linear(x, w, b, act=1), torch.nn.functional.gelu(want, approximate="tanh")))
