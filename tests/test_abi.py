"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/stableavatar_b200.h declares; argument validation answers without touching a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from stableavatar_b200 import build, _lib
    build.build()
    return _lib.lib()


def declared_symbols():
    text = (ROOT / "include" / "stableavatar_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported by libsa_b200.so"


def test_version_and_error_string(lib):
    assert lib.sa_version() >= 100
    assert isinstance(lib.sa_last_error(), bytes)


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    from stableavatar_b200 import _lib as L
    g = L.GemmArgs()                       # all-null
    assert lib.sa_gemm_bf16(C.byref(g), None) == -1
    assert b"null" in lib.sa_last_error()
    g = L.GemmArgs(a=16, w=16, out=16, M=4, N=8, K=12, lda=12, ldw=12, ldc=8)
    assert lib.sa_gemm_bf16(C.byref(g), None) == -1   # K not a multiple of 8
    a = L.AttnArgs()
    assert lib.sa_flash_attn_d128(C.byref(a), None) == -1


def test_ops_refuse_cpu_tensors():
    import torch
    from stableavatar_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_model_state_dict_keys_match_reference_names():
    """The parameter names are part of the drop-in boundary: a reference checkpoint must load."""
    from stableavatar_b200 import synth
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    cfg = synth.DIT_TINY
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    m = WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys})
    want = synth.dit_param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    assert m.config.patch_size == (1, 2, 2) and m.freqs.shape == (1024, 64) and m.freqs.dtype == torch.complex128


import torch  # noqa: E402


def test_from_pretrained_roundtrip(tmp_path):
    """Checkpoint directory contract of 1B.py:1210-1338 (config.json + safetensors, dict_mapping, in-channel padding)."""
    import json
    from safetensors.torch import save_file
    from stableavatar_b200 import synth
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    cfg = dict(synth.DIT_TINY, num_layers=1, in_dim=16)
    sd = synth.dit_state_dict(cfg)
    (tmp_path / "transformer").mkdir()
    json.dump({"in_dim": 36, "dim": 1536, "ffn_dim": cfg["ffn_dim"], "num_heads": 12, "num_layers": 1, "model_type": "i2v",
               "text_dim": cfg["text_dim"], "text_len": cfg["text_len"], "freq_dim": 256, "out_dim": 16, "eps": 1e-6,
               "unknown_key": 1}, open(tmp_path / "transformer" / "config.json", "w"))
    save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "transformer" / "diffusion_pytorch_model.safetensors"))
    m = WanTransformer3DFantasyModel.from_pretrained(str(tmp_path), subfolder="transformer",
                                                     transformer_additional_kwargs={"dict_mapping": {"in_dim": "in_channels"}},
                                                     torch_dtype=torch.bfloat16)
    assert m.dtype == torch.bfloat16 and m.in_dim == 36 and m.num_layers == 1
    w = m.patch_embedding.weight
    assert torch.equal(w[:, :16].float(), sd["patch_embedding.weight"].bfloat16().float()) and w[:, 16:].abs().max() == 0
    assert torch.equal(m.blocks[0].ffn[0].weight.float(), sd["blocks.0.ffn.0.weight"].bfloat16().float())


def test_vae_from_pretrained(tmp_path):
    from stableavatar_b200 import synth
    from stableavatar_b200.wan_vae import AutoencoderKLWan
    sd = synth.vae_state_dict()
    torch.save({k[len("model."):]: v for k, v in sd.items()}, tmp_path / "vae.pth")
    m = AutoencoderKLWan.from_pretrained(str(tmp_path / "vae.pth"), additional_kwargs={"spatial_compression_ratio": 8})
    assert torch.equal(m.state_dict()["model.decoder.conv1.weight"], sd["model.decoder.conv1.weight"])
    assert (m.config.latent_channels, m.config.temporal_compression_ratio, m.config.spacial_compression_ratio) == (16, 4, 8)


def test_riflex_table_matches_reference(golden_dir):
    """enable_riflex / disable_riflex (1B.py:891-916): the frame-axis frequency k=6 is replaced by 0.9*2*pi/L_test/scale."""
    import numpy as np
    from stableavatar_b200 import synth
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    g = np.load(golden_dir / "riflex.npz")
    cfg = dict(synth.DIT_TINY, num_layers=1)
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    m = WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys})
    base = m.freqs.clone()
    m.enable_riflex()
    assert np.allclose(m.freqs[:80].real.numpy(), g["real"], atol=1e-12) and np.allclose(m.freqs[:80].imag.numpy(), g["imag"], atol=1e-12)
    assert not torch.equal(m.freqs, base)
    m.disable_riflex()
    assert torch.equal(m.freqs, base)


def test_new_entry_points_reject_bad_arguments_without_a_gpu(lib):
    """Argument validation of the round-1 additions (sequence-parallel exchange, fused cross-attention, fp32 mode,
    VAE encode helpers) answers before any CUDA call."""
    from stableavatar_b200 import ops
    sp = ops.SpArgs()                                    # null source
    assert lib.sa_sp_scatter_qkv(C.byref(sp), None) == -1
    sp = ops.SpArgs(src=16, ld=4608, B=1, Ll=8, heads=12, head_dim=128, P=8, rank=0, hg=5)    # 5 does not divide 8
    assert lib.sa_sp_scatter_qkv(C.byref(sp), None) == -1 and b"hg" in lib.sa_last_error()
    sp = ops.SpArgs(src=16, ld=4608, B=1, Ll=8, heads=12, head_dim=128, P=2, rank=0, hg=2)    # destinations missing
    assert lib.sa_sp_scatter_o(C.byref(sp), None) == -1 and b"null destination" in lib.sa_last_error()
    assert lib.sa_sp_barrier(None, None, 2, 0, None) == -1
    ca = ops.CrossArgs()
    assert lib.sa_cross_attn3_d128(C.byref(ca), None) == -1
    ca = ops.CrossArgs(q=16, out=16, batch=1, heads=1, q_len=128, n_sets=1, scale=1.0, rows_per_group=16)
    ca.set[0] = ops.CrossSet(k=16, v=16, k_bs=1024, k_ls=128, v_bs=1024, v_ls=128, kv_len=15, kv_total=60, windowed=1)
    assert lib.sa_cross_attn3_d128(C.byref(ca), None) == -3          # (127 / 16 + 2) windows of 15 keys do not fit one step
    assert b"windowed" in lib.sa_last_error()
    assert lib.sa_f32_split3(None, C.c_int64(8), C.c_int64(4), 8, 8, None, 0, None) == -1
    assert lib.sa_f32_split3(C.c_void_p(16), C.c_int64(8), C.c_int64(4), 8, 4, C.c_void_p(16), 0, None) == -1   # K_pad < K
    assert lib.sa_f32_softmax_rows(C.c_void_p(16), 4, 8, C.c_int64(4), C.c_float(1.0), None) == -1              # ld < n
    assert lib.sa_vae_space_to_depth(C.c_void_p(16), C.c_void_p(16), 1, 5, 4, 8, None) == -1                     # odd H
    h = C.create_string_buffer(64)
    off = C.c_int64(0)
    assert lib.sa_ipc_export(None, h, C.byref(off)) == -1


def test_14b_state_dict_keys_match_reference_names():
    from stableavatar_b200 import synth
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasy14BModel
    cfg = synth.DIT_14B_TINY
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    m = WanTransformer3DFantasy14BModel(**{k: cfg[k] for k in keys})
    want = synth.dit_param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    assert "vocal_projector.proj_model.proj_2.weight" in got and "vocal_projector.proj_model.proj.weight" not in got


def test_round2_entry_points_reject_bad_arguments_without_a_gpu(lib):
    """Halo-staged conv and the attention with the fused sequence-parallel O store validate before any CUDA call."""
    from stableavatar_b200 import ops
    sup = lib.sa_conv3d_halo_supported
    assert sup(96, 96, 3, 3, 3, 1, 0) and sup(192, 192, 3, 3, 3, 1, 1) and sup(192, 384, 3, 3, 3, 1, 0) and sup(384, 192, 1, 3, 3, 1, 0)
    assert sup(96, 3, 3, 3, 3, 1, 2)                                          # the video head
    assert not sup(96, 96, 3, 1, 1, 1, 0) and not sup(32, 96, 3, 3, 3, 1, 0) and not sup(96, 96, 3, 3, 3, 2, 0)
    assert not sup(96, 128, 3, 3, 3, 1, 0) and not sup(96, 96, 3, 3, 3, 1, 3)
    a = ops.ConvArgs(inp=16, w=16, bias=16, out=16, Tout=1, H=8, W=8, Cin=32, Cout=96, KT=3, KH=3, KW=3, out_mode=0,
                     pad_h=-1, pad_w=-1, stride_t=1)
    assert lib.sa_conv3d_halo_cl(C.byref(a), None) == -3 and b"Cout 96" in lib.sa_last_error()
    assert lib.sa_conv3d_halo_cl(None, None) == -1
    from stableavatar_b200 import _lib as L
    at = L.AttnArgs()
    assert lib.sa_flash_attn_d128_sp(C.byref(at), None, 1, 1, C.c_int64(0), C.c_int64(0), None) == -1            # no destinations
    assert b"destination" in lib.sa_last_error()


def test_halo_conv_weight_packing_layout():
    """ops.pack_conv_weight_halo: [Cout, KT, KH, KW, Cin] -> [Cout / BN][taps][Cin / 8][BN][8], zero rows up to 16 for the head."""
    from stableavatar_b200 import ops
    g = torch.Generator().manual_seed(0)
    for cout, cin, kt in ((96, 48, 3), (384, 96, 1), (3, 96, 3)):
        w5 = torch.randn(cout, kt, 3, 3, cin, generator=g)
        p = ops.pack_conv_weight_halo(w5)
        bn = 16 if cout <= 16 else (96 if cout == 96 else 192)
        assert p.dtype == torch.bfloat16 and p.shape == (max(cout, bn) // bn, kt * 9, cin // 8, bn, 8)
        for (n, t, c) in ((0, 0, 0), (cout - 1, kt * 9 - 1, cin - 1), (cout // 2, 4, 17 % cin)):
            want = w5[n, t // 9, (t % 9) // 3, t % 3, c].to(torch.bfloat16)
            assert p[n // bn, t, c // 8, n % bn, c % 8] == want
        if cout < 16:
            assert (p[0, :, :, cout:] == 0).all()
