"""fp32 mode of the DiT (BASELINE config 1: fp32 weights, the reference without autocast; per-block tolerance 1e-4).

Selected automatically when the model's parameters are float32 (`model.to("cuda", torch.float32)`). Every Linear and
both attention products run on the bf16 tensor cores as split-bf16 GEMMs (csrc/fp32_kernels.cu: three bf16 terms per
operand, six cross products accumulated in the fp32 TMEM accumulator of sa_gemm_bf16), the attention softmax is a
materialised fp32 row softmax per (sample, head), and the elementwise steps that the bf16 path fuses into GEMM / norm
epilogues are separate fp32 kernels. This is the parity mode — it is not tuned for throughput, does not support
sequence parallelism or TeaCache, and is what `tests/test_fp32_gpu.py` holds against the reference's own fp32 outputs.

Mirrors 1B.py:928-1159 (forward), :650-695 (block), :383-413 / :534-605 (attentions), vp1B.py:280-450 (adapter).
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib as L
from . import ops

F32 = torch.float32


# ---------------------------------------------------------------------------------------------- C-ABI wrappers
def split3(x, pattern):
    """x f32 [M, K] (row-strided) -> bf16 [M, 6 * ceil8(K)] — sa_f32_split3."""
    assert x.dtype == F32 and x.dim() == 2 and x.stride(1) == 1 and x.is_cuda
    M, K = x.shape
    Kp = (K + 7) // 8 * 8
    out = torch.empty(M, 6 * Kp, device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().sa_f32_split3(C.c_void_p(x.data_ptr()), C.c_int64(x.stride(0)), C.c_int64(M), K, Kp,
                                  C.c_void_p(out.data_ptr()), pattern, L.stream_ptr()), "sa_f32_split3")
    return out


_W_CACHE = {}
KC = 256     # K elements per GEMM launch: the tensor core truncates when it adds into the fp32 accumulator (measured
#              ~3e-8 relative per K=16 step, same sign, so the error grows linearly with K: 1.7e-5 at K = 1536); chunks
#              of 256 are accumulated across launches by the epilogue's round-to-nearest fp32 add instead (~2e-6 total).


def _chunks(K):
    return [(k0, min(k0 + KC, K)) for k0 in range(0, K, KC)]


def split_weight(w):
    """Weight-side split of a parameter per K chunk, cached until the parameter's storage or version changes."""
    key = (w.data_ptr(), tuple(w.shape), w._version)
    hit = _W_CACHE.get(id(w))
    if hit is None or hit[0] != key:
        w2 = w.reshape(w.shape[0], -1)
        if hit is None:
            weakref.finalize(w, _W_CACHE.pop, id(w), None)       # drop the split copy with the parameter
        hit = (key, [split3(w2[:, k0:k1], 1) for k0, k1 in _chunks(w2.shape[1])])
        _W_CACHE[id(w)] = hit
    return hit[1]


def matmul_nt(x, w_chunks, bias=None, act=ops.ACT_NONE, out=None, accumulate=False):
    """out (+)= act(x @ w^T + bias) for fp32 x [M, K] and the K-chunked weight-side splits of w [N, K]."""
    N = w_chunks[0].shape[0]
    if out is None:
        out = torch.empty(x.shape[0], N, device=x.device, dtype=F32)
        accumulate = False
    ch = _chunks(x.shape[1])
    assert len(ch) == len(w_chunks)
    for i, (k0, k1) in enumerate(ch):
        last = i == len(ch) - 1
        ops.gemm(split3(x[:, k0:k1], 0), w_chunks[i], bias if i == 0 else None, act=act if last else ops.ACT_NONE, out=out,
                 res=out if (i > 0 or accumulate) else None, res_before_act=True, round_y=False)
    return out


def linear(x, w, bias=None, act=ops.ACT_NONE, w_split=None):
    """fp32 x [M, K] @ w[N, K]^T (+ bias) (+ activation) -> fp32 [M, N]."""
    return matmul_nt(x, split_weight(w) if w_split is None else w_split, bias, act)


def modulate(x, shift, scale, rows_per_batch, mod_bs):
    out = torch.empty_like(x)
    L.check(L.lib().sa_f32_modulate(C.c_void_p(x.data_ptr()), C.c_void_p(shift.data_ptr()), C.c_void_p(scale.data_ptr()),
                                    C.c_void_p(out.data_ptr()), C.c_int64(x.shape[0]), x.shape[1], rows_per_batch,
                                    C.c_int64(mod_bs), L.stream_ptr()), "sa_f32_modulate")
    return out


def gated_add_(h, y, gate=None, rows_per_batch=0, gate_bs=0):
    assert h.is_contiguous() and y.stride(1) == 1 and h.dtype == y.dtype == F32
    L.check(L.lib().sa_f32_gated_add(C.c_void_p(h.data_ptr()), C.c_void_p(y.data_ptr()), C.c_int64(y.stride(0)),
                                     C.c_void_p(L.ptr(gate)), C.c_int64(h.shape[0]), h.shape[1], rows_per_batch,
                                     C.c_int64(gate_bs), L.stream_ptr()), "sa_f32_gated_add")
    return h


def rmsnorm_rope_(x, w, freqs=None, grid=(1, 1, 1), rows_per_batch=0, eps=1e-6):
    assert x.dtype == F32 and x.dim() == 2 and x.stride(1) == 1 and w.dtype == F32
    L.check(L.lib().sa_f32_rmsnorm_rope(C.c_void_p(x.data_ptr()), C.c_int64(x.stride(0)), C.c_void_p(w.data_ptr()),
                                        C.c_void_p(L.ptr(freqs)), x.shape[0], x.shape[1], rows_per_batch, grid[0], grid[1],
                                        grid[2], C.c_float(eps), L.stream_ptr()), "sa_f32_rmsnorm_rope")
    return x


def softmax_rows_(s, scale):
    assert s.dtype == F32 and s.dim() == 2 and s.stride(1) == 1
    L.check(L.lib().sa_f32_softmax_rows(C.c_void_p(s.data_ptr()), s.shape[0], s.shape[1], C.c_int64(s.stride(0)),
                                        C.c_float(scale), L.stream_ptr()), "sa_f32_softmax_rows")
    return s


def add_bcast(a, b):
    assert a.dtype == b.dtype == F32 and a.is_contiguous() and b.is_contiguous()
    out = torch.empty(a.shape[0], b.shape[0], a.shape[1], device=a.device, dtype=F32)
    L.check(L.lib().sa_f32_add_bcast(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
                                     a.shape[0], b.shape[0], a.shape[1], L.stream_ptr()), "sa_f32_add_bcast")
    return out


def patchify(x, y, seq_len):
    B, Cx, F, H, W = x.shape
    Cy = y.shape[1] if y is not None else 0
    K_pad = ((Cx + Cy) * 4 + 7) // 8 * 8
    out = torch.empty(B, seq_len, K_pad, device=x.device, dtype=F32)
    L.check(L.lib().sa_f32_patchify(C.c_void_p(x.data_ptr()), C.c_void_p(L.ptr(y)), C.c_void_p(out.data_ptr()), B, Cx, Cy, F,
                                    H, W, seq_len, K_pad, L.stream_ptr()), "sa_f32_patchify")
    return out


def unpatchify(u, B, Cout, F, H, W):
    out = torch.empty(B, Cout, F, H, W, device=u.device, dtype=F32)
    L.check(L.lib().sa_f32_unpatchify(C.c_void_p(u.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int64(u.stride(0)),
                                      C.c_int64(u.stride(1)), B, Cout, F, H, W, L.stream_ptr()), "sa_f32_unpatchify")
    return out


def cfg_euler_step(pred, latents, dsigma, audio_scale=0.0, text_scale=0.0, cfg=True, out=None, noise_out=None, dsigma_dev=None):
    assert pred.dtype == latents.dtype == F32 and pred.is_contiguous() and latents.is_contiguous()
    n = latents.numel()
    assert pred.numel() == (3 * n if cfg else n)
    if out is None:
        out = torch.empty_like(latents)
    L.check(L.lib().sa_f32_cfg_euler_step(C.c_void_p(pred.data_ptr()), C.c_void_p(latents.data_ptr()), C.c_void_p(out.data_ptr()),
                                          C.c_void_p(L.ptr(noise_out)), C.c_int64(n), C.c_float(audio_scale), C.c_float(text_scale),
                                          C.c_float(dsigma), C.c_void_p(L.ptr(dsigma_dev)), int(cfg), L.stream_ptr()),
            "sa_f32_cfg_euler_step")
    return out


def layernorm(x, weight=None, bias=None, eps=1e-6):
    return ops.layernorm(x, weight=weight, bias=bias, eps=eps, out_dtype=F32, round_bf16=False)


def attention(q, k, v, out=None, accumulate=False):
    """softmax(q k^T / sqrt(d)) v for fp32 [B, Lq, H, D] / [B, Lk, H, D] views (last dim contiguous) -> [B, Lq, H, D].
    Per (sample, head): S = q k^T as a split GEMM into an fp32 [Lq, Lk] buffer, row softmax in place, P v as a second
    split GEMM against v^T (1B.py:158-207 with SDPA's fp32 math)."""
    B, Lq, H, D = q.shape
    Lk = k.shape[1]
    if out is None:
        out = torch.empty(B, Lq, H, D, device=q.device, dtype=F32)
        accumulate = False
    ldp = (Lk + 7) // 8 * 8
    s = torch.empty(Lq, ldp, device=q.device, dtype=F32)
    for b in range(B):
        for h in range(H):
            kh = k[b, :, h]
            matmul_nt(q[b, :, h], [split3(kh[:, k0:k1], 1) for k0, k1 in _chunks(D)], out=s[:, :Lk])
            softmax_rows_(s[:, :Lk], D ** -0.5)
            vt = v[b, :, h].t().contiguous()                                 # [D, Lk]: layout change only
            matmul_nt(s[:, :Lk], [split3(vt[:, k0:k1], 1) for k0, k1 in _chunks(Lk)], out=out[b, :, h], accumulate=accumulate)
    return out


# ---------------------------------------------------------------------------------------------- audio adapter
def adapter_forward(vp, vocal_embeddings, video_sample_n_frames, latents, e0, e):
    """FantasyTalkingVocalCondition{1B,14B}Model.forward in fp32 (vp1B.py:433-450). latents [B, L, C] f32,
    e0 [B, 6, C] f32, e [B, C] f32 -> ([B, G, A, Ca] f32, lens)."""
    B, T, _ = vocal_embeddings.shape
    Ca = vp.audio_proj_dim
    pm = vp.proj_model
    a = vocal_embeddings.reshape(B * T, -1).to(F32).contiguous()
    if hasattr(pm, "proj_1"):
        feat = layernorm(linear(a, pm.proj_1.weight), pm.norm_1.weight, pm.norm_1.bias, 1e-5)
        feat = layernorm(linear(feat, pm.proj_2.weight), pm.norm_2.weight, pm.norm_2.bias, 1e-5)
    else:
        feat = layernorm(linear(a, pm.proj.weight), pm.norm.weight, pm.norm.bias, 1e-5)
    table, G, A, lens = vp._window_table(T, video_sample_n_frames, feat.device)
    Lt = latents.shape[1]
    if Lt % G != 0:
        raise RuntimeError(f"shape '[{B * G}, -1, 8, {Ca // 8}]' is invalid for input of size {B * Lt * Ca}")
    mods = torch.stack([blk.modulation.reshape(-1) for blk in vp.blocks]).contiguous()          # [2, 6Ca]
    outs = []
    for b in range(B):
        x = ops.gather_rows(feat[b * T:(b + 1) * T].contiguous(), table.reshape(-1))            # [G*A, Ca]
        eb = add_bcast(mods, e0[b].reshape(1, -1).contiguous())                                 # [2, 1, 6Ca]
        lat = latents[b]
        rows = G * A
        for i, blk in enumerate(vp.blocks):
            ch = eb[i, 0].view(6, Ca)
            ca = blk.cross_attn
            gated_add_(x, modulate(layernorm(x), ch[0], ch[1], rows, 0), ch[2], rows, 0)       # pseudo self-attention
            xn = layernorm(x, blk.norm3.weight, blk.norm3.bias)
            q = rmsnorm_rope_(linear(xn, ca.q.weight, ca.q.bias), ca.norm_q.weight)
            kk = rmsnorm_rope_(linear(lat, ca.k.weight, ca.k.bias), ca.norm_k.weight)
            vv = linear(lat, ca.v.weight, ca.v.bias)
            nh, hd = ca.num_heads, ca.head_dim
            at = attention(q.view(G, A, nh, hd), kk.view(G, Lt // G, nh, hd), vv.view(G, Lt // G, nh, hd))
            gated_add_(x, linear(at.view(rows, Ca), ca.o.weight, ca.o.bias))
            xm = modulate(layernorm(x), ch[3], ch[4], rows, 0)
            y = linear(linear(xm, blk.ffn[0].weight, blk.ffn[0].bias, act=ops.ACT_GELU_TANH), blk.ffn[2].weight, blk.ffn[2].bias)
            gated_add_(x, y, ch[5], rows, 0)
        em = add_bcast(vp.final_head.modulation.reshape(2, -1).contiguous(), e[b].reshape(1, -1).contiguous())   # [2, 1, Ca]
        xf = modulate(layernorm(x), em[0, 0], em[1, 0], rows, 0)
        fh = vp.final_head.final_proj
        outs.append(linear(xf, fh.weight, fh.bias).view(G, A, Ca))
    ctx = torch.stack(outs)
    if B > 1:
        lens = torch.cat([lens] * 3)
    return ctx, lens


# ---------------------------------------------------------------------------------------------- DiT forward
@torch.no_grad()
def forward(model, x, t, context, seq_len, clip_fea=None, y=None, cond_flag=True, vocal_embeddings=None,
            is_clip_level_modeling=False, video_sample_n_frames=81):
    if model.sp_world_size > 1 or model.teacache is not None:
        raise NotImplementedError("fp32 mode is the single-GPU parity mode: no sequence parallelism, no TeaCache")
    dev = model.device
    C_, nh = model.dim, model.num_heads
    if isinstance(x, (list, tuple)):
        x = torch.stack(list(x))
    if isinstance(y, (list, tuple)):
        y = torch.stack(list(y))
    B, _, F, H, W = x.shape
    Hp, Wp = H // 2, W // 2
    Lv = F * Hp * Wp
    assert Lv <= seq_len
    Lt = seq_len

    A = patchify(x.to(dev, F32).contiguous(), None if y is None else y.to(dev, F32).contiguous(), Lt)
    K = model.in_dim * 4
    w_pe = model.patch_embedding.weight.reshape(C_, K)
    if A.shape[-1] != K:                                                     # zero-padded K columns of the operand gather
        w_pad = torch.zeros(C_, A.shape[-1], device=dev, dtype=F32)
        w_pad[:, :K] = w_pe
        w_pe = w_pad
    h = linear(A.view(B * Lt, -1), w_pe, model.patch_embedding.bias, w_split=[split3(w_pe.contiguous(), 1)])
    if Lt > Lv:
        h.view(B, Lt, C_)[:, Lv:].zero_()

    te, tp = model.time_embedding, model.time_projection[1]
    t32 = t.to(device=dev, dtype=F32).contiguous()
    h1, _ = ops.small_linear(t32, te[0].weight, te[0].bias, pre=2)
    e32, _ = ops.small_linear(h1, te[2].weight, te[2].bias, pre=1)
    e0, _ = ops.small_linear(e32, tp.weight, tp.bias, pre=1)                                     # [B, 6C] f32

    ctx_in = torch.zeros(B, model.text_len, model.text_dim, device=dev, dtype=F32)
    for i, u in enumerate(context):
        ctx_in[i, :u.size(0)] = u
    txe = model.text_embedding
    ctx_txt = linear(linear(ctx_in.view(B * model.text_len, -1), txe[0].weight, txe[0].bias, act=ops.ACT_GELU_TANH),
                     txe[2].weight, txe[2].bias)
    ip = model.img_emb.proj
    n_img = clip_fea.shape[1]
    if n_img != 257:
        raise RuntimeError("cross-attention expects 257 CLIP tokens (context[:, :257], 1B.py:544)")
    c = layernorm(clip_fea.to(dev, F32).reshape(B * n_img, -1).contiguous(), ip[0].weight, ip[0].bias, 1e-5)
    c = linear(linear(c, ip[1].weight, ip[1].bias, act=ops.ACT_GELU_ERF), ip[3].weight, ip[3].bias)
    ctx_img = layernorm(c, ip[4].weight, ip[4].bias, 1e-5)

    h3, e0_3 = h.view(B, Lt, C_), e0.view(B, 6, C_)
    vocal_embeddings = vocal_embeddings.to(dev, F32)
    vp = model.vocal_projector
    if vocal_embeddings.size(0) > 1 and model._cfg_audio_trick:
        vc, _ = adapter_forward(vp, vocal_embeddings[-1:], video_sample_n_frames, h3[-1:], e0_3[-1:], e32[-1:])
        vc = torch.cat([torch.zeros_like(vc), vc, vc])
    else:
        vc, _ = adapter_forward(vp, vocal_embeddings, video_sample_n_frames, h3, e0_3, e32)
    G = (video_sample_n_frames - 1) // 4 + 1
    if vc.shape[0] != B:
        raise RuntimeError(f"audio context batch {vc.shape[0]} != latent batch {B}")
    grouped = not is_clip_level_modeling
    if model.hooks is not None:
        model.hooks["vocal_context"] = vc.flatten(1, 2) if is_clip_level_modeling else vc
        model.hooks["e0"] = e0_3
    vc2 = vc.reshape(B, -1, C_).contiguous()

    mods = torch.stack([blk.modulation.reshape(-1) for blk in model.blocks]).contiguous()
    e_all = add_bcast(mods, e0)                                                                  # [layers, B, 6C]
    freqs = model._freqs_table(dev)
    for i, blk in enumerate(model.blocks):
        e = e_all[i]
        ch = [e[:, k * C_:(k + 1) * C_] for k in range(6)]
        sa, ca = blk.self_attn, blk.cross_attn
        # self-attention (1B.py:383-413)
        t1 = modulate(layernorm(h), ch[0], ch[1], Lt, 6 * C_)
        q = rmsnorm_rope_(linear(t1, sa.q.weight, sa.q.bias), sa.norm_q.weight, freqs, (F, Hp, Wp), Lt)
        k = rmsnorm_rope_(linear(t1, sa.k.weight, sa.k.bias), sa.norm_k.weight, freqs, (F, Hp, Wp), Lt)
        v = linear(t1, sa.v.weight, sa.v.bias)
        a = attention(q.view(B, Lt, nh, 128), k.view(B, Lt, nh, 128), v.view(B, Lt, nh, 128))
        gated_add_(h, linear(a.view(B * Lt, C_), sa.o.weight, sa.o.bias), ch[2], Lt, 6 * C_)
        # cross-attention: text + CLIP image + audio share q (1B.py:534-605)
        xn = layernorm(h, blk.norm3.weight, blk.norm3.bias)
        q = rmsnorm_rope_(linear(xn, ca.q.weight, ca.q.bias), ca.norm_q.weight).view(B, Lt, nh, 128)
        kt = rmsnorm_rope_(linear(ctx_txt, ca.k.weight, ca.k.bias), ca.norm_k.weight).view(B, -1, nh, 128)
        vt = linear(ctx_txt, ca.v.weight, ca.v.bias).view(B, -1, nh, 128)
        ki = rmsnorm_rope_(linear(ctx_img, ca.k_img.weight, ca.k_img.bias), ca.norm_k_img.weight).view(B, -1, nh, 128)
        vi = linear(ctx_img, ca.v_img.weight, ca.v_img.bias).view(B, -1, nh, 128)
        kv_ = linear(vc2.view(-1, C_), ca.k_vocal.weight, ca.k_vocal.bias)
        vv_ = linear(vc2.view(-1, C_), ca.v_vocal.weight, ca.v_vocal.bias)
        a = attention(q, ki, vi)
        attention(q, kt, vt, out=a, accumulate=True)
        if grouped:
            if Lt % G != 0:
                raise RuntimeError(f"shape '[{B * G}, -1, {nh}, 128]' is invalid for input of size {B * Lt * C_}")
            attention(q.view(B * G, Lt // G, nh, 128), kv_.view(B * G, -1, nh, 128), vv_.view(B * G, -1, nh, 128),
                      out=a.view(B * G, Lt // G, nh, 128), accumulate=True)
        else:
            attention(q, kv_.view(B, -1, nh, 128), vv_.view(B, -1, nh, 128), out=a, accumulate=True)
        gated_add_(h, linear(a.view(B * Lt, C_), ca.o.weight, ca.o.bias))
        # FFN (1B.py:687-691)
        t2 = modulate(layernorm(h), ch[3], ch[4], Lt, 6 * C_)
        y2 = linear(linear(t2, blk.ffn[0].weight, blk.ffn[0].bias, act=ops.ACT_GELU_TANH), blk.ffn[2].weight, blk.ffn[2].bias)
        gated_add_(h, y2, ch[5], Lt, 6 * C_)
        if model.hooks is not None:
            model.hooks[f"block{i}"] = h.view(B, Lt, C_).clone()

    em = add_bcast(model.head.modulation.reshape(2, -1).contiguous(), e32)                       # [2, B, C]
    xh = modulate(layernorm(h), em[0], em[1], Lt, C_)
    u = linear(xh, model.head.head.weight, model.head.head.bias).view(B, Lt, -1)
    return unpatchify(u, B, model.out_dim, F, H, W)
