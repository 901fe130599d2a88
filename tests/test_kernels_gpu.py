"""GPU tests (-m gpu) of the individual C-ABI kernels: against torch fp32 references at small / ragged sizes, and at the
full BASELINE sizes (L = 32 760 tokens, C = 1536, ffn 8960) through size-independent properties where a CPU oracle
would take minutes: constant-V attention returns the constant, key-permutation invariance, GEMM linearity, norm
statistics."""
import pytest
import torch

pytestmark = pytest.mark.gpu
L_FULL, C, H = 32760, 1536, 12


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from stableavatar_b200 import ops
    return ops


def sdpa(q, k, v):
    f = torch.nn.functional.scaled_dot_product_attention
    return f(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float()).transpose(1, 2)


@pytest.mark.parametrize("B,Lq,Lk,heads", [(1, 256, 256, 1), (2, 300, 333, 3), (3, 130, 64, 2), (2, 77, 1, 2), (63, 156, 15, 12),
                                           (1, 1000, 769, 4)])
def test_flash_attention_vs_torch(ops, B, Lq, Lk, heads):
    g = torch.Generator(device="cuda").manual_seed(Lq * 7 + Lk)
    q, k, v = (torch.randn(B, n, heads, 128, device="cuda", generator=g).bfloat16() for n in (Lq, Lk, Lk))
    out = ops.flash_attn(q, k, v)
    assert rel(out, sdpa(q, k, v)) < 5e-3
    base = torch.randn(B, Lq, heads, 128, device="cuda", generator=g).bfloat16()
    acc = ops.flash_attn(q, k, v, out=base.clone(), accumulate=True)
    assert rel(acc, base.float() + sdpa(q, k, v).bfloat16().float()) < 5e-3


def test_flash_attention_strided_qkv_views(ops):
    """q/k/v as views of one fused [B*L, 3C] projection buffer (how the DiT calls it)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    B, L, nh = 2, 384, 3
    qkv = torch.randn(B * L, 3 * nh * 128, device="cuda", generator=g).bfloat16()
    v5 = qkv.view(B, L, 3, nh, 128)
    out = ops.flash_attn(v5[:, :, 0], v5[:, :, 1], v5[:, :, 2])
    assert rel(out, sdpa(v5[:, :, 0], v5[:, :, 1], v5[:, :, 2])) < 5e-3


@pytest.mark.parametrize("B,Lq,heads,n_dst,all_heads", [(2, 300, 3, 2, 5), (1, 1000, 2, 4, 2), (3, 130, 2, 1, 4), (1, 640, 1, 5, 3)])
def test_flash_attention_row_partitioned_destinations(ops, B, Lq, heads, n_dst, all_heads):
    """sa_flash_attn_d128_sp (the sequence-parallel O exchange fused into the epilogue) with local destinations: query row r
    lands in dst[r // rows] at [b, r % rows, head_offset + head], bit-identical to the ordinary call; tiles that straddle two
    destinations (negative TMA start row) and ragged last destinations included; untouched head columns stay untouched."""
    g = torch.Generator(device="cuda").manual_seed(Lq + n_dst)
    q, k, v = (torch.randn(B, n, heads, 128, device="cuda", generator=g).bfloat16() for n in (Lq, 333, 333))
    want = ops.flash_attn(q, k, v)
    rows = -(-Lq // n_dst)
    h0 = all_heads - heads                         # this "rank" owns the last `heads` head columns of every destination
    dst = [torch.full((B, rows, all_heads, 128), 7.0, device="cuda").bfloat16() for _ in range(n_dst)]
    ops.flash_attn_sp(q, k, v, [d.data_ptr() + h0 * 128 * 2 for d in dst], rows, rows * all_heads * 128, all_heads * 128)
    torch.cuda.synchronize()
    for j, d in enumerate(dst):
        n = min(rows, Lq - j * rows)
        assert torch.equal(d[:, :n, h0:], want[:, j * rows:j * rows + n])
        assert (d[:, :, :h0] == 7.0).all() and (d[:, n:, h0:] == 7.0).all()


def test_flash_attention_large_logits_rescale_path(ops):
    """Scores that grow by far more than 2^8 along the key axis force the lazy O / l rescale."""
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(1, 256, 2, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 512, 2, 128, device="cuda", generator=g)
    k = (k * torch.linspace(0.2, 6.0, 512, device="cuda").view(1, -1, 1, 1)).bfloat16()
    v = torch.randn(1, 512, 2, 128, device="cuda", generator=g).bfloat16()
    assert rel(ops.flash_attn(q, k, v), sdpa(q, k, v)) < 1e-2


def test_full_length_attention_properties(ops):
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(1, L_FULL, 2, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, L_FULL, 2, 128, device="cuda", generator=g).bfloat16()
    const = torch.randn(128, device="cuda", generator=g).bfloat16()
    out = ops.flash_attn(q, k, const.expand(1, L_FULL, 2, 128).contiguous())
    assert (out.float() - const.float()).abs().max().item() <= 2e-2 * const.float().abs().max().item()   # rows of P sum to 1
    v = torch.randn(1, L_FULL, 2, 128, device="cuda", generator=g).bfloat16()
    a = ops.flash_attn(q, k, v)
    perm = torch.randperm(L_FULL, device="cuda", generator=g)
    b = ops.flash_attn(q, k[:, perm].contiguous(), v[:, perm].contiguous())
    assert rel(a, b) < 5e-3                                                                              # key order is irrelevant
    sub = slice(0, 512)
    assert rel(a[:, sub], sdpa(q[:, sub], k, v)) < 5e-3


@pytest.mark.parametrize("M,N,K", [(256, 512, 256), (1000, 200, 136), (777, 1536, 1536), (5, 64, 1536), (300, 4608, 144)])
def test_gemm_epilogues_vs_torch(ops, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g).bfloat16()
    y = (a.float() @ w.float().t() + bias.float()).bfloat16().float()
    assert rel(ops.gemm(a, w, bias), y) < 4e-3
    assert rel(ops.gemm(a, w, bias, act=ops.ACT_GELU_TANH), torch.nn.functional.gelu(y, approximate="tanh")) < 6e-3
    assert rel(ops.gemm(a, w, bias, act=ops.ACT_GELU_ERF), torch.nn.functional.gelu(y)) < 6e-3
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    assert rel(ops.gemm(a, w, bias, res=res), res.float() + y) < 4e-3
    gate = torch.randn(3, N, device="cuda", generator=g).bfloat16()
    rpb = (M + 2) // 3
    gi = torch.arange(M, device="cuda") // rpb
    want = res.float() + (y * gate.float()[gi]).bfloat16().float()
    assert rel(ops.gemm(a, w, bias, res=res.clone(), gate=gate, gate_ld=N, rows_per_batch=rpb), want) < 4e-3
    # fp32 output / fp32 residual (adapter stream), no rounding of the Linear output
    r32 = torch.randn(M, N, device="cuda", generator=g)
    out = ops.gemm(a, w, bias, res=r32, out=torch.empty(M, N, device="cuda"), round_y=True)
    assert rel(out, r32 + y) < 2e-4                              # a few 1-ulp bf16 roundings of y differ from torch's accumulation order


def test_full_size_gemm_linearity(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    M = 3 * L_FULL
    a = torch.randn(M, C, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(8960, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    w2 = (torch.randn(8960, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    y1 = ops.gemm(a, w1, out_dtype=torch.bfloat16).float()
    y12 = ops.gemm(a, torch.cat([w1, w2]))                       # N = 17920: more column tiles, same values
    assert torch.equal(y12[:, :8960].float(), y1)
    rows = torch.randint(0, M, (256,), device="cuda", generator=g)
    want = a[rows].float() @ w2.float().t()
    assert rel(y12[rows, 8960:], want) < 4e-3


@pytest.mark.parametrize("C", [1536, 2048, 5120])        # 1.3B width, the widest 8-chunk row, the 14B width (20 chunks)
def test_layernorm_modulate_and_rmsnorm_rope_vs_torch(ops, C):
    H = C // 128
    from oracle import dit as O
    g = torch.Generator(device="cuda").manual_seed(2)
    B, F, Hh, W = 2, 3, 5, 7
    L = F * Hh * W + 9                                           # 9 zero-padded (un-rotated) tokens
    x = torch.randn(B * L, C, device="cuda", generator=g).bfloat16()
    e = (torch.randn(B, 6 * C, device="cuda", generator=g) * 0.3).bfloat16()
    out = ops.layernorm(x, shift=e[:, :C], scale=e[:, C:2 * C], mod_bs=6 * C, rows_per_batch=L)
    xn = torch.nn.functional.layer_norm(x.float(), (C,), eps=1e-6).bfloat16()
    want = (xn * (1 + e[:, C:2 * C]).repeat_interleave(L, 0) + e[:, :C].repeat_interleave(L, 0))
    assert rel(out, want.float()) < 3e-3
    w = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).bfloat16()
    bb = (0.1 * torch.randn(C, device="cuda", generator=g)).bfloat16()
    out = ops.layernorm(x, weight=w, bias=bb, out_dtype=torch.float32, round_bf16=False, eps=1e-5)
    assert rel(out, torch.nn.functional.layer_norm(x.float(), (C,), w.float(), bb.float(), 1e-5)) < 1e-5
    # RMSNorm + RoPE against the oracle's fp64 rope_apply
    q = torch.randn(B * L, C, device="cuda", generator=g).bfloat16()
    fr = O.rope_freqs(128)
    table = torch.stack([fr.real, fr.imag], -1).float().contiguous().cuda()
    got = ops.rmsnorm_rope_(q.clone(), w, freqs=table, grid=(F, Hh, W), rows_per_batch=L)
    qn = O.rms_norm(q.float().cpu().view(B, L, C).bfloat16(), w.cpu()).view(B, L, H, 128)
    want = O.rope_apply(qn, [(F, Hh, W)] * B, fr).view(B * L, C)
    assert rel(got.cpu(), want) < 4e-3


def test_full_size_norm_statistics(ops):
    g = torch.Generator(device="cuda").manual_seed(4)
    x = (torch.randn(3 * L_FULL, C, device="cuda", generator=g) * 3 + 1).bfloat16()
    y = ops.layernorm(x).float()
    assert y.mean(-1).abs().max().item() < 1e-2 and (y.std(-1, unbiased=False) - 1).abs().max().item() < 1e-2
    w = torch.ones(C, device="cuda").bfloat16()
    r = ops.rmsnorm_rope_(x.clone(), w).float()
    assert ((r * r).mean(-1) - 1).abs().max().item() < 2e-2


def test_config1_grid_480x832x5_vs_oracle():
    """BASELINE config 1 grid (480x832, 5 frames: L = 2 x 30 x 52 = 3120) on the 2-layer stand-in model, CFG batch 3."""
    from oracle import dit as O
    from stableavatar_b200 import synth
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    cfg = synth.DIT_TINY
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    sd = {k: v.bfloat16() for k, v in synth.dit_state_dict(cfg).items()}
    m = WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys})
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda", torch.bfloat16)
    inp = synth.dit_inputs(cfg, frames=5, height=480, width=832, seed=7)
    bf, dev = torch.bfloat16, "cuda"
    out = m(x=inp["x"].to(dev, bf), t=inp["t"].to(dev), context=[c.to(dev, bf) for c in inp["context"]], seq_len=inp["seq_len"],
            clip_fea=inp["clip_fea"].to(dev, bf), y=inp["y"].to(dev, bf), vocal_embeddings=inp["vocal_embeddings"].to(dev, bf),
            video_sample_n_frames=5)
    r = lambda t: t.bfloat16().float()  # noqa: E731
    with torch.no_grad():
        ref = O.dit_forward({k: v.float() for k, v in sd.items()}, cfg, r(inp["x"]), inp["t"], [r(c) for c in inp["context"]],
                            inp["seq_len"], r(inp["clip_fea"]), r(inp["y"]), r(inp["vocal_embeddings"]), 5)
    assert out.shape == (3, 16, 2, 60, 104)
    assert rel(out.float().cpu(), ref) < 2e-2


def _cross_inputs(B, Lq, Hh, G, A, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()  # noqa: E731
    return mk(B, Lq, Hh, 128), (mk(B, 512, Hh, 128), mk(B, 512, Hh, 128)), (mk(B, 257, Hh, 128), mk(B, 257, Hh, 128)), \
        (mk(B, G * A, Hh, 128), mk(B, G * A, Hh, 128))


@pytest.mark.parametrize("Lq,G,tok_offset,L_total", [(3120, 2, 0, 3120),      # config-1 grid: 2 windows of 1560 rows
                                                     (1700, 3, 1000, 4680),   # token shard [1000, 2700) of 3 x 1560
                                                     (520, 4, 0, 520)])       # 130-row groups: 4 windows inside one CTA
def test_fused_cross_attention_equals_three_launches(ops, Lq, G, tok_offset, L_total):
    """sa_cross_attn3_d128 (text + image + windowed audio in one launch) against three sa_flash_attn_d128 launches with
    accumulate and against fp32 torch SDPA per set. The plain sets are bit-identical by construction (next test); the
    windowed step of the fused kernel keeps every exponential on MUFU (where a window sits inside the step depends on the
    token sharding, and the sequence-parallel forward must stay bit-identical), while a separate 15-key launch takes a few
    of them from the FMA-pipe polynomial (relative error 7.5e-5 before the bf16 rounding of P)."""
    B, Hh, A = 2, 3, 15
    gs = L_total // G
    q, (kt, vt), (ki, vi), (ka, va) = _cross_inputs(B, Lq, Hh, G, A, 11)
    got = ops.cross_attn3(q, [(kt, vt, 0), (ki, vi, 0), (ka, va, A)], rows_per_group=gs, tok_offset=tok_offset)
    want = ops.flash_attn(q, kt, vt)
    ops.flash_attn(q, ki, vi, out=want, accumulate=True)
    ref = sdpa(q, kt, vt) + sdpa(q, ki, vi)
    for g in range((tok_offset) // gs, (tok_offset + Lq - 1) // gs + 1):
        lo, hi = max(g * gs, tok_offset) - tok_offset, min((g + 1) * gs, tok_offset + Lq) - tok_offset
        ops.flash_attn(q[:, lo:hi], ka[:, g * A:(g + 1) * A], va[:, g * A:(g + 1) * A], out=want[:, lo:hi], accumulate=True)
        ref[:, lo:hi] += sdpa(q[:, lo:hi], ka[:, g * A:(g + 1) * A], va[:, g * A:(g + 1) * A])
    torch.cuda.synchronize()
    assert rel(got.float(), want.float()) < 1.5e-3 and (got.float() - want.float()).abs().max() <= 2.0 ** -6 * want.float().abs().max()
    assert rel(got.float(), ref) < 6e-3


def test_fused_cross_attention_plain_sets_and_accumulate(ops):
    """Clip-level audio (no windows): three plain sets; accumulate adds onto an existing output."""
    q, (kt, vt), (ki, vi), (ka, va) = _cross_inputs(1, 300, 2, 3, 15, 12)
    base = torch.randn(q.shape, device="cuda").bfloat16()
    got = ops.cross_attn3(q, [(kt, vt, 0), (ki, vi, 0), (ka, va, 0)], out=base.clone(), accumulate=True)
    want = base.clone()
    for k, v in ((kt, vt), (ki, vi), (ka, va)):
        ops.flash_attn(q, k, v, out=want, accumulate=True)
    assert torch.equal(got, want)
    one = ops.cross_attn3(q, [(ki, vi, 0)])
    assert torch.equal(one, ops.flash_attn(q, ki, vi))


@pytest.mark.parametrize("cin,cout,T,H,W,mode,with_res", [(96, 96, 2, 16, 8, 0, False), (96, 96, 3, 37, 21, 0, True),
                                                          (192, 192, 2, 40, 24, 0, True), (96, 192, 1, 16, 16, 1, False),
                                                          (192, 96, 2, 33, 50, 0, False), (384, 192, 1, 20, 12, 0, True),
                                                          (192, 384, 2, 24, 16, 0, False), (384, 384, 1, 15, 26, 0, True)])
@pytest.mark.parametrize("kt", [3, 1])
def test_halo_conv_vs_torch_conv3d(ops, cin, cout, T, H, W, mode, with_res, kt):
    """sa_conv3d_halo_cl (input halo staged once, taps as descriptor offsets, weights shared by 2-4 tiles) against
    torch conv3d in fp32 on the same bf16 operands, and against the per-tap kernel sa_conv3d_cl it replaces: ragged tiles,
    residual add, the frame-interleaved output of the temporal upsampler, 1-4 channel groups."""
    g = torch.Generator(device="cuda").manual_seed(cin + cout + H)
    x = torch.randn(T + kt - 1, H, W, cin, device="cuda", generator=g).bfloat16()
    w5 = (torch.randn(cout, kt, 3, 3, cin, device="cuda", generator=g) * (9 * kt * cin) ** -0.5).bfloat16()
    bias = torch.randn(cout, device="cuda", generator=g)
    oshape = (2 * T, H, W, cout // 2) if mode == 1 else (T, H, W, cout)
    res = torch.randn(oshape, device="cuda", generator=g).bfloat16() if with_res else None
    assert ops.conv3d_halo_supported(cin, cout, (kt, 3, 3), 1, mode)
    out = ops.conv3d_halo_cl(x, ops.pack_conv_weight_halo(w5), bias, cout=cout, out=torch.empty(oshape, device="cuda", dtype=torch.bfloat16),
                             res=res, out_mode=mode, kt=kt)
    old = ops.conv3d_cl(x, w5.reshape(cout, -1).contiguous(), bias, cout=cout, k=(kt, 3, 3),
                        out=torch.empty(oshape, device="cuda", dtype=torch.bfloat16), res=res, out_mode=mode)
    ref = torch.nn.functional.conv3d(x.float().permute(3, 0, 1, 2)[None], w5.float().permute(0, 4, 1, 2, 3), bias,
                                     padding=(0, 1, 1))[0].permute(1, 2, 3, 0)                       # [T, H, W, cout]
    if mode == 1:
        ref = torch.stack([ref[..., :cout // 2], ref[..., cout // 2:]], dim=1).reshape(oshape)
    if with_res:
        ref = ref + res.float()
    assert rel(out, ref) < 4e-3 and rel(old, ref) < 4e-3
    assert rel(out, old) < 3e-3


def test_halo_conv_video_head_vs_per_tap_kernel(ops):
    """Decoder head (96 -> 3, fp32 planar clamped output into a frame window of a longer video) on the halo kernel."""
    g = torch.Generator(device="cuda").manual_seed(11)
    T, H, W, cin, cout = 2, 40, 27, 96, 3
    x = torch.randn(T + 2, H, W, cin, device="cuda", generator=g).bfloat16()
    w5 = (torch.randn(cout, 3, 3, 3, cin, device="cuda", generator=g) * (27 * cin) ** -0.5 * 2).bfloat16()
    bias = torch.randn(cout, device="cuda", generator=g) * 0.3
    assert ops.conv3d_halo_supported(cin, cout, (3, 3, 3), 1, 2)
    new = torch.full((cout, 5, H, W), 9.0, device="cuda")
    old = torch.full((cout, 5, H, W), 9.0, device="cuda")
    ops.conv3d_halo_cl(x, ops.pack_conv_weight_halo(w5), bias, cout=cout, out=new, out_mode=2, out_T_total=5, out_t0=2)
    w16 = torch.zeros(16, 27 * cin, device="cuda", dtype=torch.bfloat16)
    w16[:cout] = w5.reshape(cout, -1)
    ops.conv3d_cl(x, w16, bias, cout=cout, k=(3, 3, 3), out=old, out_mode=2, out_T_total=5, out_t0=2)
    assert (new[:, :2] == 9.0).all() and (new[:, 4:] == 9.0).all()
    assert (new[:, 2:4].abs() <= 1.0).all() and (new[:, 2:4].abs() == 1.0).any()           # the clamp is exercised
    assert rel(new[:, 2:4], old[:, 2:4]) < 3e-3
