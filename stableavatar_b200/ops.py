"""Host-side wrappers over the C-ABI (include/stableavatar_b200.h). torch tensors are used as device memory only:
each wrapper fills the argument struct from pointers/strides and launches on the current CUDA stream. There is no
eager fallback — a missing library or a CPU tensor raises."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._lib import ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, ACT_SILU  # noqa: F401


TIMING = None  # set to {} to collect (start, end) CUDA-event pairs per tag on the launching stream (bench.py)


class timed:
    """with ops.timed("self_attn"): ... — brackets the enclosed launches with CUDA events when ops.TIMING is a dict."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if TIMING is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            TIMING.setdefault(self.tag, []).append((self.e0, e1))


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("stableavatar_b200 ops need CUDA tensors (no CPU fallback)")


def gemm(a, w, bias=None, *, out=None, act=ACT_NONE, res=None, gate=None, gate_ld=0, rows_per_batch=0, round_y=True,
         out_dtype=torch.bfloat16, res_before_act=False):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T) — see sa_gemm_bf16. a may be a row-strided 2-D view."""
    _need_cuda(a, w)
    assert a.dim() == 2 and w.dim() == 2 and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    assert out.stride(1) == 1 and out.shape == (M, N)
    g = L.GemmArgs(a=a.data_ptr(), w=w.data_ptr(), out=out.data_ptr(), bias=L.ptr(bias), res=L.ptr(res),
                   gate=L.ptr(gate), lda=a.stride(0), ldw=w.stride(0), ldc=out.stride(0),
                   ldr=res.stride(0) if res is not None else 0, gate_ld=gate_ld, M=M, N=N, K=K,
                   bias_dtype=L.dt(bias) if bias is not None else 0, out_dtype=L.dt(out),
                   res_dtype=L.dt(res) if res is not None else 0, act=act,
                   res_mode=0 if res is None else (3 if res_before_act else (2 if gate is not None else 1)), round_y=int(round_y),
                   rows_per_batch=rows_per_batch)
    L.check(L.lib().sa_gemm_bf16(C.byref(g), L.stream_ptr()), "sa_gemm_bf16")
    return out


def _attn_args(q, k, v, out, accumulate, scale):
    B, Lq, H, D = q.shape
    assert k.shape[0] == B and k.shape[2] == H and k.shape[3] == D and v.shape == k.shape
    for t in (q, k, v, out):
        assert t.dtype == torch.bfloat16 and t.stride(3) == 1 and t.stride(2) == D
    return L.AttnArgs(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), out=out.data_ptr(), q_bs=q.stride(0),
                      q_ls=q.stride(1), k_bs=k.stride(0), k_ls=k.stride(1), v_bs=v.stride(0), v_ls=v.stride(1),
                      o_bs=out.stride(0), o_ls=out.stride(1), batch=B, heads=H, q_len=Lq, kv_len=k.shape[1],
                      scale=scale if scale is not None else D ** -0.5, accumulate=int(accumulate))


def flash_attn(q, k, v, out=None, accumulate=False, scale=None):
    """softmax(q k^T * scale) v for [B, L, H, 128] bf16 views (token/batch strides free) — sa_flash_attn_d128."""
    _need_cuda(q, k, v)
    if q.shape[3] != 128:
        raise NotImplementedError("flash_attn: head_dim 128 only")
    if out is None:
        assert not accumulate
        out = torch.empty(q.shape, device=q.device, dtype=torch.bfloat16)
    g = _attn_args(q, k, v, out, accumulate, scale)
    L.check(L.lib().sa_flash_attn_d128(C.byref(g), L.stream_ptr()), "sa_flash_attn_d128")
    return out


def flash_attn_sp(q, k, v, dst_ptrs, rows_per_dst, dst_bs, dst_ls, scale=None):
    """flash_attn whose epilogue stores query row r straight into dst_ptrs[r // rows_per_dst] at [b, r % rows_per_dst, head]
    (device addresses, typically peers' o_recv over NVLink) — sa_flash_attn_d128_sp, the fused sequence-parallel O exchange."""
    _need_cuda(q, k, v)
    if q.shape[3] != 128:
        raise NotImplementedError("flash_attn: head_dim 128 only")
    g = _attn_args(q, k, v, q, False, scale)                 # `out` is ignored by the entry point
    n = len(dst_ptrs)
    arr = (C.c_void_p * 8)(*(list(dst_ptrs) + [None] * (8 - n)))
    L.check(L.lib().sa_flash_attn_d128_sp(C.byref(g), arr, n, rows_per_dst, C.c_int64(dst_bs), C.c_int64(dst_ls),
                                          L.stream_ptr()), "sa_flash_attn_d128_sp")


def attn_small_q(q, k, v, out=None, scale=None):
    """Few-queries attention (audio adapter), any head_dim % 8 == 0 — sa_attn_small_q."""
    _need_cuda(q, k, v)
    if out is None:
        out = torch.empty(q.shape, device=q.device, dtype=torch.bfloat16)
    g = _attn_args(q, k, v, out, False, scale)
    L.check(L.lib().sa_attn_small_q(C.byref(g), C.c_int32(q.shape[3]), L.stream_ptr()), "sa_attn_small_q")
    return out


class LnArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("shift", C.c_void_p), ("scale", C.c_void_p), ("gate", C.c_void_p), ("res", C.c_void_p),
                ("ldx", C.c_int64), ("ldo", C.c_int64), ("ldr", C.c_int64), ("mod_bs", C.c_int64),
                ("rows", C.c_int32), ("C", C.c_int32), ("rows_per_batch", C.c_int32),
                ("x_dtype", C.c_int32), ("out_dtype", C.c_int32), ("w_dtype", C.c_int32), ("round_bf16", C.c_int32),
                ("eps", C.c_float)]


def layernorm(x, *, weight=None, bias=None, shift=None, scale=None, gate=None, res=None, mod_bs=0, rows_per_batch=0,
              eps=1e-6, out=None, out_dtype=torch.bfloat16, round_bf16=True):
    """LayerNorm (+affine) (+modulate) (+gated residual) over the last dim of a 2-D row view — sa_layernorm_modulate."""
    _need_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    rows, Cc = x.shape
    if out is None:
        out = torch.empty(rows, Cc, device=x.device, dtype=out_dtype)
    for t in (shift, scale, gate):
        assert t is None or t.dtype == torch.bfloat16
    if res is not None:
        assert res.dtype == out.dtype
    a = LnArgs(x=x.data_ptr(), out=out.data_ptr(), weight=L.ptr(weight), bias=L.ptr(bias), shift=L.ptr(shift),
               scale=L.ptr(scale), gate=L.ptr(gate), res=L.ptr(res), ldx=x.stride(0), ldo=out.stride(0),
               ldr=res.stride(0) if res is not None else 0, mod_bs=mod_bs, rows=rows, C=Cc,
               rows_per_batch=rows_per_batch, x_dtype=L.dt(x), out_dtype=L.dt(out),
               w_dtype=L.dt(weight) if weight is not None else 0, round_bf16=int(round_bf16), eps=eps)
    L.check(L.lib().sa_layernorm_modulate(C.byref(a), L.stream_ptr()), "sa_layernorm_modulate")
    return out


class RmsArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("weight", C.c_void_p), ("x2", C.c_void_p), ("weight2", C.c_void_p),
                ("freqs", C.c_void_p), ("ld", C.c_int64), ("rows", C.c_int32), ("C", C.c_int32),
                ("rows_per_batch", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("tok_offset", C.c_int32),
                ("eps", C.c_float)]


def rmsnorm_rope_(x, weight, x2=None, weight2=None, freqs=None, grid=(1, 1, 1), rows_per_batch=0, eps=1e-6,
                  tok_offset=0):
    """In-place RMSNorm (+RoPE) on one or two bf16 row views sharing a row stride — sa_rmsnorm_rope."""
    _need_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
    if x2 is not None:
        assert x2.shape == x.shape and x2.stride() == x.stride() and weight2.dtype == torch.bfloat16
    if freqs is not None:
        assert freqs.dtype == torch.float32 and freqs.shape == (1024, 64, 2) and freqs.is_contiguous()
    a = RmsArgs(x=x.data_ptr(), weight=weight.data_ptr(), x2=L.ptr(x2), weight2=L.ptr(weight2), freqs=L.ptr(freqs),
                ld=x.stride(0), rows=x.shape[0], C=x.shape[1], rows_per_batch=rows_per_batch, F=grid[0], H=grid[1],
                W=grid[2], tok_offset=tok_offset, eps=eps)
    L.check(L.lib().sa_rmsnorm_rope(C.byref(a), L.stream_ptr()), "sa_rmsnorm_rope")
    return x


def add_bcast(a, b):
    """out[i, j, :] = bf16(a[i, :] + b[j, :])."""
    _need_cuda(a, b)
    assert a.dtype == b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous() and a.shape[1] == b.shape[1]
    out = torch.empty(a.shape[0], b.shape[0], a.shape[1], device=a.device, dtype=torch.bfloat16)
    L.check(L.lib().sa_add_bcast_bf16(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
                                      a.shape[0], b.shape[0], a.shape[1], L.stream_ptr()), "sa_add_bcast_bf16")
    return out


def patchify(x, y, seq_len):
    _need_cuda(x)
    B, Cx, F, H, W = x.shape
    Cy = y.shape[1] if y is not None else 0
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and (y is None or (y.dtype == torch.bfloat16 and y.is_contiguous()))
    K = (Cx + Cy) * 4
    K_pad = (K + 7) // 8 * 8
    out = torch.empty(B, seq_len, K_pad, device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().sa_patchify(C.c_void_p(x.data_ptr()), C.c_void_p(L.ptr(y)), C.c_void_p(out.data_ptr()), B, Cx, Cy,
                                F, H, W, seq_len, K_pad, L.stream_ptr()), "sa_patchify")
    return out


def unpatchify(u, B, Cout, F, H, W):
    _need_cuda(u)
    assert u.dim() == 3 and u.dtype == torch.bfloat16 and u.stride(2) == 1
    out = torch.empty(B, Cout, F, H, W, device=u.device, dtype=torch.bfloat16)
    L.check(L.lib().sa_unpatchify(C.c_void_p(u.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int64(u.stride(0)),
                                  C.c_int64(u.stride(1)), B, Cout, F, H, W, L.stream_ptr()), "sa_unpatchify")
    return out


def small_linear(x, w, bias, pre=0, want_f32=True, want_bf16=False, K=None):
    """fp32 rows: out = pre(x) @ w^T + bias — sa_small_linear_f32 (8 rows per launch; a batch of W window triples has
    3W rows). pre=2: x is t[M], K = sinusoid width."""
    _need_cuda(x, w)
    assert x.dtype == torch.float32 and x.is_contiguous() and w.is_contiguous()
    M = x.shape[0]
    N, Kw = w.shape
    of = torch.empty(M, N, device=x.device, dtype=torch.float32) if want_f32 else None
    ob = torch.empty(M, N, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    for m0 in range(0, M, 8):
        m = min(8, M - m0)
        L.check(L.lib().sa_small_linear_f32(C.c_void_p(x[m0:].data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(L.ptr(bias)),
                                            C.c_void_p(None if of is None else of[m0:].data_ptr()),
                                            C.c_void_p(None if ob is None else ob[m0:].data_ptr()), m, N, Kw, pre, L.dt(w),
                                            L.stream_ptr()), "sa_small_linear_f32")
    return of, ob


def gather_rows(src, idx):
    _need_cuda(src, idx)
    assert src.dtype == torch.float32 and src.is_contiguous() and idx.dtype == torch.int32
    out = torch.empty(idx.numel(), src.shape[1], device=src.device, dtype=torch.float32)
    L.check(L.lib().sa_gather_rows_f32(C.c_void_p(src.data_ptr()), C.c_void_p(idx.data_ptr()), C.c_void_p(out.data_ptr()),
                                       idx.numel(), src.shape[1], L.stream_ptr()), "sa_gather_rows_f32")
    return out


def cfg_euler_step(pred, latents, dsigma, audio_scale=0.0, text_scale=0.0, cfg=True, out=None, noise_out=None,
                   dsigma_dev=None):
    """latents' = bf16(float(latents) + dsigma * CFG(pred)) — sa_cfg_euler_step. latents: bf16, or fp32 for a
    caller-supplied fp32 sample; the result is bf16 (the model output's dtype)."""
    _need_cuda(pred, latents)
    assert pred.dtype == torch.bfloat16 and latents.dtype in (torch.bfloat16, torch.float32)
    assert pred.is_contiguous() and latents.is_contiguous()
    n = latents.numel()
    assert pred.numel() == (3 * n if cfg else n)
    if out is None:
        out = torch.empty(latents.shape, device=latents.device, dtype=torch.bfloat16)
    L.check(L.lib().sa_cfg_euler_step(C.c_void_p(pred.data_ptr()), C.c_void_p(latents.data_ptr()),
                                      C.c_void_p(out.data_ptr()), C.c_void_p(L.ptr(noise_out)), C.c_int64(n),
                                      C.c_float(audio_scale), C.c_float(text_scale), C.c_float(dsigma), C.c_void_p(L.ptr(dsigma_dev)), int(cfg),
                                      L.dt(latents), L.stream_ptr()), "sa_cfg_euler_step")
    return out


class WindowBlendArgs(C.Structure):
    _fields_ = [("new_latents", C.c_void_p), ("pred_latents", C.c_void_p), ("n_windows", C.c_int32), ("C", C.c_int32),
                ("N", C.c_int32), ("HW", C.c_int32), ("f_max", C.c_int32), ("overlap", C.c_int32), ("pred_dtype", C.c_int32),
                ("start", C.c_int32 * 64), ("frames", C.c_int32 * 64), ("prev_end", C.c_int32 * 64),
                ("blend", C.c_int32 * 64), ("weight", C.c_float * 64), ("one_minus_weight", C.c_float * 64)]


def window_blend_(pred_latents, new_latents, windows, overlap, weight, one_minus_weight):
    """Write all windows of a step back into pred_latents [1, C, N, h, w] with the overlap blend, in window order —
    sa_window_blend. new_latents: bf16 [W, C, f_max, h, w]; windows: (start, frames, prev_end, blend) per window;
    weight / one_minus_weight: the bf16 tensor values of the reference's weight tensor and of (1 - weight)."""
    _need_cuda(pred_latents, new_latents)
    assert new_latents.dtype == torch.bfloat16 and new_latents.is_contiguous() and pred_latents.is_contiguous()
    assert pred_latents.dim() == 5 and pred_latents.shape[0] == 1 and new_latents.dim() == 5
    W, Cc, f_max, h, w = new_latents.shape
    assert len(windows) == W and pred_latents.shape[1] == Cc and tuple(pred_latents.shape[3:]) == (h, w)
    a = WindowBlendArgs(new_latents=new_latents.data_ptr(), pred_latents=pred_latents.data_ptr(), n_windows=W, C=Cc,
                        N=pred_latents.shape[2], HW=h * w, f_max=f_max, overlap=overlap, pred_dtype=L.dt(pred_latents))
    for k, (ws, f, prev_end, blend) in enumerate(windows):
        a.start[k], a.frames[k], a.prev_end[k], a.blend[k] = ws, f, prev_end, int(blend)
    for j in range(overlap):
        a.weight[j], a.one_minus_weight[j] = float(weight[j]), float(one_minus_weight[j])
    L.check(L.lib().sa_window_blend(C.byref(a), L.stream_ptr()), "sa_window_blend")
    return pred_latents


# ---------------------------------------------------------------------------------------------- Wan VAE ops
class ConvArgs(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("res", C.c_void_p), ("out", C.c_void_p),
                ("Tout", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
                ("KT", C.c_int32), ("KH", C.c_int32), ("KW", C.c_int32),
                ("out_mode", C.c_int32), ("out_T_total", C.c_int32), ("out_t0", C.c_int32),
                ("pad_h", C.c_int32), ("pad_w", C.c_int32), ("stride_t", C.c_int32)]


def conv3d_cl(x, w, bias, *, cout, k, out, res=None, out_mode=0, out_T_total=0, out_t0=0, pad=None, stride_t=1):
    """Causal conv on channels-last bf16 x [(Tout-1)*stride_t + KT, H, W, Cin] (leading frames = cache) — sa_conv3d_cl.
    pad = (front pad H, front pad W), default 'same' (K // 2)."""
    _need_cuda(x, w)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and w.dtype == torch.bfloat16 and w.is_contiguous()
    assert bias.dtype == torch.float32 and out.is_contiguous() and (res is None or (res.is_contiguous() and res.dtype == torch.bfloat16))
    Tin, H, W, Cin = x.shape
    a = ConvArgs(inp=x.data_ptr(), w=w.data_ptr(), bias=bias.data_ptr(), res=L.ptr(res), out=out.data_ptr(),
                 Tout=(Tin - k[0]) // stride_t + 1, H=H, W=W, Cin=Cin, Cout=cout, KT=k[0], KH=k[1], KW=k[2],
                 out_mode=out_mode, out_T_total=out_T_total, out_t0=out_t0, pad_h=-1 if pad is None else pad[0],
                 pad_w=-1 if pad is None else pad[1], stride_t=stride_t)
    L.check(L.lib().sa_conv3d_cl(C.byref(a), L.stream_ptr()), "sa_conv3d_cl")
    return out


def conv3d_halo_supported(cin, cout, k, stride_t=1, out_mode=0):
    return bool(L.lib().sa_conv3d_halo_supported(cin, cout, k[0], k[1], k[2], stride_t, out_mode))


def pack_conv_weight_halo(w5):
    """[Cout, KT, KH, KW, Cin] -> bf16 [Cout / BN][taps][Cin / 8][BN][8] (BN = 96 or 192 output channels per tile), the tcgen05
    no-swizzle K-major layout sa_conv3d_halo_cl reads."""
    cout, kt, kh, kw, cin = w5.shape
    if cout <= 16:                                     # the video head: one 16-channel tile, rows >= cout zero
        w5 = torch.cat([w5, w5.new_zeros(16 - cout, kt, kh, kw, cin)]) if cout < 16 else w5
        cout = 16
    bn = cout if cout in (16, 96) else 192
    return w5.reshape(cout // bn, bn, kt * kh * kw, cin // 8, 8).permute(0, 2, 3, 1, 4).to(torch.bfloat16).contiguous()


def conv3d_halo_cl(x, w_packed, bias, *, cout, out, res=None, out_mode=0, kt=3, out_T_total=0, out_t0=0):
    """(kt)x3x3 'same' causal conv on channels-last bf16 x [Tout + kt - 1, H, W, Cin] with halo staging — sa_conv3d_halo_cl."""
    _need_cuda(x, w_packed)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and w_packed.dtype == torch.bfloat16 and w_packed.is_contiguous()
    assert bias.dtype == torch.float32 and out.is_contiguous() and (res is None or (res.is_contiguous() and res.dtype == torch.bfloat16))
    Tin, H, W, Cin = x.shape
    a = ConvArgs(inp=x.data_ptr(), w=w_packed.data_ptr(), bias=bias.data_ptr(), res=L.ptr(res), out=out.data_ptr(),
                 Tout=Tin - (kt - 1), H=H, W=W, Cin=Cin, Cout=cout, KT=kt, KH=3, KW=3, out_mode=out_mode, out_T_total=out_T_total, out_t0=out_t0,
                 pad_h=-1, pad_w=-1, stride_t=1)
    L.check(L.lib().sa_conv3d_halo_cl(C.byref(a), L.stream_ptr()), "sa_conv3d_halo_cl")
    return out


def vae_rmsnorm_silu(x, gamma, out, silu=True):
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and out.is_contiguous() and gamma.dtype == torch.float32
    Cc = x.shape[-1]
    L.check(L.lib().sa_vae_rmsnorm_silu(C.c_void_p(x.data_ptr()), C.c_void_p(gamma.data_ptr()), C.c_void_p(out.data_ptr()),
                                        C.c_int64(x.numel() // Cc), Cc, int(silu), L.stream_ptr()), "sa_vae_rmsnorm_silu")
    return out


def vae_upsample2x(x, out):
    _need_cuda(x)
    T, H, W, Cc = x.shape
    assert x.is_contiguous() and out.is_contiguous() and out.shape == (T, 2 * H, 2 * W, Cc)
    L.check(L.lib().sa_vae_upsample2x(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), T, H, W, Cc, L.stream_ptr()),
            "sa_vae_upsample2x")
    return out


def softmax_rows(s, scale):
    _need_cuda(s)
    assert s.dtype == torch.float32 and s.dim() == 2 and s.stride(1) == 1
    out = torch.empty(s.shape, device=s.device, dtype=torch.bfloat16)
    L.check(L.lib().sa_softmax_rows(C.c_void_p(s.data_ptr()), C.c_void_p(out.data_ptr()), s.shape[0], s.shape[1],
                                    C.c_int64(s.stride(0)), C.c_int64(out.stride(0)), C.c_float(scale), L.stream_ptr()),
            "sa_softmax_rows")
    return out


def softmax_rows_into(s, out, scale):
    """out[r, :] = bf16(softmax(s[r, :] * scale)) for row-strided 2-D views."""
    _need_cuda(s, out)
    assert s.dtype == torch.float32 and out.dtype == torch.bfloat16 and s.shape == out.shape and s.stride(1) == out.stride(1) == 1
    L.check(L.lib().sa_softmax_rows(C.c_void_p(s.data_ptr()), C.c_void_p(out.data_ptr()), s.shape[0], s.shape[1],
                                    C.c_int64(s.stride(0)), C.c_int64(out.stride(0)), C.c_float(scale), L.stream_ptr()),
            "sa_softmax_rows")
    return out


def vae_latent_in(z, wc, bc, mean, std, cpad):
    _need_cuda(z)
    Cz = z.shape[0]
    P = z.numel() // Cz
    assert z.dtype == torch.float32 and z.is_contiguous()
    out = torch.empty(*z.shape[1:], cpad, device=z.device, dtype=torch.bfloat16)
    L.check(L.lib().sa_vae_latent_in(C.c_void_p(z.data_ptr()), C.c_void_p(wc.data_ptr()), C.c_void_p(bc.data_ptr()),
                                     C.c_void_p(mean.data_ptr()), C.c_void_p(std.data_ptr()), C.c_void_p(out.data_ptr()),
                                     Cz, C.c_int64(P), cpad, L.stream_ptr()), "sa_vae_latent_in")
    return out


def vae_space_to_depth(x, out):
    """bf16 [T, H, W, C] -> [T, H/2, W/2, 4C] with channel (dy*2 + dx)*C + c — sa_vae_space_to_depth."""
    _need_cuda(x)
    T, H, W, Cc = x.shape
    assert x.is_contiguous() and out.is_contiguous() and out.shape == (T, H // 2, W // 2, 4 * Cc) and x.dtype == out.dtype == torch.bfloat16
    L.check(L.lib().sa_vae_space_to_depth(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), T, H, W, Cc, L.stream_ptr()),
            "sa_vae_space_to_depth")
    return out


def vae_video_in(x, cpad):
    """f32 planar [Cx, T, H, W] -> bf16 channels-last [T, H, W, cpad] — sa_vae_video_in."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    Cx = x.shape[0]
    out = torch.empty(*x.shape[1:], cpad, device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().sa_vae_video_in(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), Cx, C.c_int64(x.numel() // Cx), cpad,
                                    L.stream_ptr()), "sa_vae_video_in")
    return out


def vae_latent_out(h, wc, bc, mean, std, out):
    """f32 channels-last h [T, H, W, 2Cz] -> f32 planar out [2Cz, T, H, W] (conv1 + latent normalisation) — sa_vae_latent_out."""
    _need_cuda(h)
    C2 = h.shape[-1]
    assert h.dtype == torch.float32 and h.is_contiguous() and out.dtype == torch.float32 and out.is_contiguous()
    assert out.numel() == h.numel() and wc.numel() == C2 * C2
    L.check(L.lib().sa_vae_latent_out(C.c_void_p(h.data_ptr()), C.c_void_p(wc.data_ptr()), C.c_void_p(bc.data_ptr()),
                                      C.c_void_p(mean.data_ptr()), C.c_void_p(std.data_ptr()), C.c_void_p(out.data_ptr()),
                                      C2 // 2, C.c_int64(h.numel() // C2), L.stream_ptr()), "sa_vae_latent_out")
    return out


# ---------------------------------------------------------------------------------------------- sequence-parallel exchange
class SpArgs(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst_a", C.c_void_p * 8), ("dst_b", C.c_void_p * 8), ("ld", C.c_int64),
                ("B", C.c_int32), ("Ll", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32), ("P", C.c_int32),
                ("rank", C.c_int32), ("hg", C.c_int32), ("b_first", C.c_int32), ("b_count", C.c_int32)]


def _sp_args(src, dst_a, dst_b, ld, B, Ll, heads, P, rank, hg, b_first=0, b_count=0):
    a = SpArgs(src=src.data_ptr(), ld=ld, B=B, Ll=Ll, heads=heads, head_dim=128, P=P, rank=rank, hg=hg, b_first=b_first,
               b_count=b_count)
    for r in range(P):
        a.dst_a[r] = dst_a[r]
        a.dst_b[r] = dst_b[r] if dst_b is not None else None
    return a


def sp_scatter_qkv(qkv, kv_ptrs, q_ptrs, *, B, Ll, heads, P, rank, hg, b_first=0, b_count=0):
    """qkv: local [B*Ll, >= 3*heads*128] bf16 rows (q | k | v) -> peers' kv_recv / q_recv — sa_sp_scatter_qkv.
    b_first / b_count: the CFG samples moved by this launch (0 = all)."""
    _need_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.dim() == 2 and qkv.stride(1) == 1 and qkv.shape[0] == B * Ll
    a = _sp_args(qkv, kv_ptrs, q_ptrs, qkv.stride(0), B, Ll, heads, P, rank, hg, b_first, b_count)
    L.check(L.lib().sa_sp_scatter_qkv(C.byref(a), L.stream_ptr()), "sa_sp_scatter_qkv")


def sp_scatter_o(o, o_ptrs, *, B, Ll, heads, P, rank, hg, b_first=0, b_count=0):
    """o: this rank's attention output [B, P/qs * Ll, heads/hg, 128] bf16 contiguous -> peers' o_recv — sa_sp_scatter_o."""
    _need_cuda(o)
    assert o.dtype == torch.bfloat16 and o.is_contiguous()
    a = _sp_args(o, o_ptrs, None, 0, B, Ll, heads, P, rank, hg, b_first, b_count)
    L.check(L.lib().sa_sp_scatter_o(C.byref(a), L.stream_ptr()), "sa_sp_scatter_o")


def sp_barrier(sig_ptrs, epoch, P, rank):
    """Flag barrier across the ranks whose flag arrays are mapped at sig_ptrs — sa_sp_barrier."""
    arr = (C.c_void_p * 8)(*([sig_ptrs[r] for r in range(P)] + [None] * (8 - P)))
    L.check(L.lib().sa_sp_barrier(arr, C.c_void_p(epoch.data_ptr()), P, rank, L.stream_ptr()), "sa_sp_barrier")


def sp_set_barrier_timeout_ms(ms: int):
    """Spin bound of sa_sp_barrier (default 10 min; 0 = unbounded like NCCL) — sa_sp_set_barrier_timeout_ms."""
    rc = L.lib().sa_sp_set_barrier_timeout_ms(C.c_int64(int(ms)))
    if rc != 0:
        raise RuntimeError(f"sa_sp_set_barrier_timeout_ms failed (code {rc}): {L.lib().sa_last_error().decode()}")


def ipc_export(t):
    """(64-byte handle, byte offset) of the device allocation holding tensor t — sa_ipc_export."""
    _need_cuda(t)
    h = C.create_string_buffer(64)
    off = C.c_int64(0)
    L.check(L.lib().sa_ipc_export(C.c_void_p(t.data_ptr()), h, C.byref(off)), "sa_ipc_export")
    return h.raw, off.value


def ipc_open(handle):
    """Base address of a peer process's allocation mapped for the CURRENT device — sa_ipc_open."""
    base = C.c_void_p(0)
    L.check(L.lib().sa_ipc_open(C.c_char_p(handle), C.byref(base)), "sa_ipc_open")
    return base.value


def ipc_close(base):
    L.check(L.lib().sa_ipc_close(C.c_void_p(base)), "sa_ipc_close")


# ---------------------------------------------------------------------------------------------- fused cross-attention
class CrossSet(C.Structure):
    _fields_ = [("k", C.c_void_p), ("v", C.c_void_p), ("k_bs", C.c_int64), ("k_ls", C.c_int64), ("v_bs", C.c_int64),
                ("v_ls", C.c_int64), ("kv_len", C.c_int32), ("kv_total", C.c_int32), ("windowed", C.c_int32)]


class CrossArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("out", C.c_void_p), ("q_bs", C.c_int64), ("q_ls", C.c_int64), ("o_bs", C.c_int64),
                ("o_ls", C.c_int64), ("batch", C.c_int32), ("heads", C.c_int32), ("q_len", C.c_int32), ("n_sets", C.c_int32),
                ("scale", C.c_float), ("accumulate", C.c_int32), ("rows_per_group", C.c_int32), ("tok_offset", C.c_int32),
                ("set", CrossSet * 3)]


def cross_attn3(q, sets, out=None, accumulate=False, rows_per_group=0, tok_offset=0, scale=None):
    """out (+)= sum_s softmax(q K_s^T * scale) V_s in one launch — sa_cross_attn3_d128. q/out [B, Lq, H, 128] bf16 views;
    sets: up to three (k, v, window) with k, v [B, Lk, H, 128] views; window = 0 for a plain set, else the keys are
    Lk / window consecutive windows and row r attends to window (tok_offset + r) // rows_per_group."""
    _need_cuda(q)
    B, Lq, H, D = q.shape
    assert D == 128 and 1 <= len(sets) <= 3
    if out is None:
        assert not accumulate
        out = torch.empty(q.shape, device=q.device, dtype=torch.bfloat16)
    for t in (q, out):
        assert t.dtype == torch.bfloat16 and t.stride(3) == 1 and t.stride(2) == D
    a = CrossArgs(q=q.data_ptr(), out=out.data_ptr(), q_bs=q.stride(0), q_ls=q.stride(1), o_bs=out.stride(0), o_ls=out.stride(1),
                  batch=B, heads=H, q_len=Lq, n_sets=len(sets), scale=scale if scale is not None else D ** -0.5,
                  accumulate=int(accumulate), rows_per_group=rows_per_group, tok_offset=tok_offset)
    for i, (k, v, window) in enumerate(sets):
        assert k.shape == v.shape and k.shape[0] == B and k.shape[2] == H and k.shape[3] == D
        for t in (k, v):
            assert t.dtype == torch.bfloat16 and t.stride(3) == 1 and t.stride(2) == D
        a.set[i] = CrossSet(k=k.data_ptr(), v=v.data_ptr(), k_bs=k.stride(0), k_ls=k.stride(1), v_bs=v.stride(0), v_ls=v.stride(1),
                            kv_len=window if window else k.shape[1], kv_total=k.shape[1], windowed=int(bool(window)))
    L.check(L.lib().sa_cross_attn3_d128(C.byref(a), L.stream_ptr()), "sa_cross_attn3_d128")
    return out


def sp_norm_rope_scatter(qkv, weight_q, weight_k, kv_ptrs, q_ptrs, *, B, Ll, heads, P, rank, hg, freqs=None, grid=(1, 1, 1),
                         tok_offset=0, eps=1e-6, b_first=0, b_count=0):
    """RMSNorm + RoPE of the q / k parts of local QKV rows fused with the peer scatter — sa_sp_norm_rope_scatter."""
    _need_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.dim() == 2 and qkv.stride(1) == 1 and qkv.shape[0] == B * Ll
    assert weight_q.dtype == weight_k.dtype == torch.bfloat16
    if freqs is not None:
        assert freqs.dtype == torch.float32 and freqs.shape == (1024, 64, 2) and freqs.is_contiguous()
    a = _sp_args(qkv, kv_ptrs, q_ptrs, qkv.stride(0), B, Ll, heads, P, rank, hg, b_first, b_count)
    L.check(L.lib().sa_sp_norm_rope_scatter(C.byref(a), C.c_void_p(weight_q.data_ptr()), C.c_void_p(weight_k.data_ptr()),
                                            C.c_void_p(L.ptr(freqs)), grid[0], grid[1], grid[2], tok_offset, C.c_float(eps),
                                            L.stream_ptr()), "sa_sp_norm_rope_scatter")
