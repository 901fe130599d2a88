#!/usr/bin/env python
"""Library-kernel baseline (NOT product, not on any product path): the same denoise step — 1.3B audio-DiT, CFG batch 3,
L = 32 760 — built from stock PyTorch ops in the reference's op order under bf16 autocast, i.e. what the reference's own
GPU path dispatches to on a B200: F.scaled_dot_product_attention (cuDNN / flash SDPA), F.linear (cuBLASLt),
F.layer_norm, elementwise kernels, complex128 RoPE (wan/models/wan_fantasy_transformer3d_1B.py:296-323, 383-413,
534-605, 650-695, 928-1159). BASELINE.md §3 asks for this number beside ours. Self-contained on purpose (it imports
neither the product package's kernels nor oracle/): weights are random, only the timing matters.

    python tools/lib_baseline.py --steps 5 --warmup 2 [--out profiles/r02_lib_baseline.json]
"""
from __future__ import annotations

import argparse
import json
import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent

DIM, FFN, HEADS, LAYERS, TEXT_LEN, TEXT_DIM, FREQ = 1536, 8960, 12, 30, 512, 4096, 256


def rope_params(n, dim, theta=10000):
    f = torch.outer(torch.arange(n), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(f), f)


def rope_apply(x, grid, freqs):
    """1B.py:296-323: per-sample complex128 rotation, fp32 result."""
    n, c = x.size(2), x.size(3) // 2
    freqs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    out = []
    f, h, w = grid
    for i in range(x.size(0)):
        seq = f * h * w
        xi = torch.view_as_complex(x[i, :seq].to(torch.float64).reshape(seq, n, -1, 2))
        fi = torch.cat([freqs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), freqs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                        freqs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq, 1, -1)
        xi = torch.view_as_real(xi * fi).flatten(2)
        out.append(torch.cat([xi, x[i, seq:]]))
    return torch.stack(out).float()


def rms(x, w, eps=1e-6):
    return (x.float() * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * w


def ln(x, w=None, b=None, eps=1e-6):
    return F.layer_norm(x.float(), (x.shape[-1],), None if w is None else w.float(), None if b is None else b.float(), eps).type_as(x)


def sdpa(q, k, v):
    """attention(), 1B.py:158-207 SDPA branch: [B, L, N, D] in / out."""
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    return o.transpose(1, 2).contiguous()


class Block(torch.nn.Module):
    def __init__(s):
        super().__init__()
        L = torch.nn.Linear
        s.q, s.k, s.v, s.o = L(DIM, DIM), L(DIM, DIM), L(DIM, DIM), L(DIM, DIM)
        s.cq, s.ck, s.cv, s.co, s.ck_img, s.cv_img, s.ck_voc, s.cv_voc = (L(DIM, DIM) for _ in range(8))
        s.f0, s.f2 = L(DIM, FFN), L(FFN, DIM)
        s.nq, s.nk, s.cnq, s.cnk, s.cnk_img = (torch.nn.Parameter(torch.ones(DIM)) for _ in range(5))
        s.n3w, s.n3b = torch.nn.Parameter(torch.ones(DIM)), torch.nn.Parameter(torch.zeros(DIM))
        s.mod = torch.nn.Parameter(torch.randn(1, 6, DIM) / DIM ** 0.5)

    def forward(s, x, e0, grid, freqs, ctx, vc, G):
        B, L, C = x.shape
        n, d = HEADS, C // HEADS
        e = (s.mod + e0).chunk(6, dim=1)
        h = ln(x) * (1 + e[1]) + e[0]
        q = rms(s.q(h), s.nq).view(B, L, n, d)
        k = rms(s.k(h), s.nk).view(B, L, n, d)
        v = s.v(h).view(B, L, n, d)
        a = sdpa(rope_apply(q, grid, freqs).to(v.dtype), rope_apply(k, grid, freqs).to(v.dtype), v)
        x = x + s.o(a.flatten(2)) * e[2]
        xn = ln(x, s.n3w, s.n3b)
        img, txt = ctx[:, :257], ctx[:, 257:]
        q = rms(s.cq(xn), s.cnq).view(B, -1, n, d)
        a = sdpa(q, rms(s.ck(txt), s.cnk).view(B, -1, n, d), s.cv(txt).view(B, -1, n, d))
        a = a + sdpa(q, rms(s.ck_img(img), s.cnk_img).view(B, -1, n, d), s.cv_img(img).view(B, -1, n, d))
        kv, vv = s.ck_voc(vc).view(B * G, -1, n, d), s.cv_voc(vc).view(B * G, -1, n, d)
        a = a + sdpa(q.view(B * G, -1, n, d), kv, vv).view(B, -1, n, d)
        x = x + s.co(a.flatten(2))
        h = ln(x) * (1 + e[4]) + e[3]
        return x + s.f2(F.gelu(s.f0(h), approximate="tanh")) * e[5]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    Fl, Hp, Wp = 21, 30, 52
    L, B, G = Fl * Hp * Wp, 3, 21
    with torch.device(dev):
        blocks = torch.nn.ModuleList([Block() for _ in range(LAYERS)]).to(torch.bfloat16)
        patch = torch.nn.Conv3d(36, DIM, (1, 2, 2), (1, 2, 2)).to(torch.bfloat16)
        text = torch.nn.Sequential(torch.nn.Linear(TEXT_DIM, DIM), torch.nn.GELU(approximate="tanh"), torch.nn.Linear(DIM, DIM)).to(torch.bfloat16)
        timee = torch.nn.Sequential(torch.nn.Linear(FREQ, DIM), torch.nn.SiLU(), torch.nn.Linear(DIM, DIM)).to(torch.bfloat16)
        timep = torch.nn.Sequential(torch.nn.SiLU(), torch.nn.Linear(DIM, DIM * 6)).to(torch.bfloat16)
        head = torch.nn.Linear(DIM, 64).to(torch.bfloat16)
        x_in = torch.randn(B, 36, Fl, 60, 104, dtype=torch.bfloat16)
        ctx_in = torch.randn(B, TEXT_LEN, TEXT_DIM, dtype=torch.bfloat16) * 0.1
        img_ctx = torch.randn(B, 257, DIM, dtype=torch.bfloat16)
        vc = torch.randn(B, G * 15, DIM, dtype=torch.bfloat16)
        t_emb = torch.randn(B, FREQ)
    d = DIM // HEADS
    freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)), rope_params(1024, 2 * (d // 6))], dim=1).to(dev)

    @torch.no_grad()
    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            x = patch(x_in).flatten(2).transpose(1, 2)
            e0 = timep(timee(t_emb.to(torch.bfloat16))).unflatten(1, (6, DIM))      # three tiny GEMMs (fp32 island in the reference)
            ctx = torch.cat([img_ctx, text(ctx_in)], dim=1)
            for b in blocks:
                x = b(x, e0, (Fl, Hp, Wp), freqs, ctx, vc, G)
            u = head(ln(x))
            pred = u.view(B, Fl, Hp, Wp, 1, 2, 2, 16).permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, 16, Fl, 60, 104)
            uu, dd, cc = pred.chunk(3)
            noise = uu + 5.0 * (dd - uu) + 3.0 * (cc - dd)
            return (x_in[:1, :16].float() + (-0.02) * noise).to(torch.bfloat16)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    s = e0.elapsed_time(e1) / 1e3 / args.steps

    # the attention call alone (what flash_attn_v8 replaces), same shapes
    q = torch.randn(B, L, HEADS, d, device=dev, dtype=torch.bfloat16)
    k, v = torch.randn_like(q), torch.randn_like(q)
    for _ in range(3):
        sdpa(q, k, v)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(10):
        sdpa(q, k, v)
    a1.record()
    torch.cuda.synchronize()
    attn_ms = a0.elapsed_time(a1) / 10
    backends = {n: getattr(torch.backends.cuda, n)() for n in ("flash_sdp_enabled", "mem_efficient_sdp_enabled", "cudnn_sdp_enabled",
                                                               "math_sdp_enabled") if hasattr(torch.backends.cuda, n)}
    res = {"what": "stock-PyTorch (library kernels) denoise step, 1.3B audio-DiT, 480x832x81f, CFG batch 3, bf16 autocast, eager",
           "library_step_s": s, "steps": args.steps, "warmup": args.warmup, "sdpa_ms_per_call_B3": attn_ms,
           "sdpa_tflops": 4.0 * B * L * L * DIM / (attn_ms * 1e-3) / 1e12, "torch": torch.__version__, "sdp_backends": backends,
           "gpu": torch.cuda.get_device_name(0),
           "note": "adapter (0.6 TFLOP of 855) and the CLIP MLP are left out of this baseline; RoPE in complex128 as the reference"}
    print(json.dumps(res))
    if args.out:
        Path(args.out).write_text(json.dumps(res, indent=1) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
