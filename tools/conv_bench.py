"""Per-tap conv kernel (sa_conv3d_cl) vs halo-staged kernel (sa_conv3d_halo_cl) at the Wan VAE decoder's dominant shapes:
one chunk of 4 frames at 480x832 (96 -> 96) and at 240x416 (192 -> 192); rotation in one process, CUDA events."""
import sys

import torch

sys.path.insert(0, ".")
from stableavatar_b200 import ops  # noqa: E402

PEAK = 1391.5
SHAPES = [(96, 96, 4, 480, 832, 3), (192, 192, 4, 240, 416, 3), (96, 192, 4, 240, 416, 3), (192, 96, 4, 480, 832, 3),
          (384, 384, 2, 120, 208, 3), (192, 384, 2, 120, 208, 3), (384, 384, 1, 60, 104, 3), (192, 96, 4, 480, 832, 1),
          (384, 192, 4, 240, 416, 1), (384, 192, 2, 120, 208, 1)]
for cin, cout, T, H, W, kt in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(T + kt - 1, H, W, cin, device="cuda", generator=g).bfloat16()
    w5 = (torch.randn(cout, kt, 3, 3, cin, device="cuda", generator=g) * (9 * kt * cin) ** -0.5).bfloat16()
    bias = torch.randn(cout, device="cuda", generator=g)
    wp, wo = ops.pack_conv_weight_halo(w5), w5.reshape(cout, -1).contiguous()
    out = torch.empty(T, H, W, cout, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * T * H * W * 9 * kt * cin * cout

    def run(fn, iters=10):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    old = lambda: ops.conv3d_cl(x, wo, bias, cout=cout, k=(kt, 3, 3), out=out)  # noqa: E731
    new = lambda: ops.conv3d_halo_cl(x, wp, bias, cout=cout, out=out, kt=kt)  # noqa: E731
    ts = {"per-tap": [], "halo": []}
    for _ in range(4):
        ts["per-tap"].append(run(old))
        ts["halo"].append(run(new))
    for k, v in ts.items():
        m = sorted(v)[len(v) // 2]
        print(f"{cin:3d}->{cout:3d} k{kt} {T}x{H}x{W} {k:8s}: {m:.3f} ms = {fl / m / 1e9:.0f} TFLOP/s = {fl / m / 1e9 / PEAK:.2f} of sustained peak", flush=True)
