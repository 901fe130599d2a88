"""Deterministic synthetic weights and inputs: there are no checkpoints or datasets on the build or
GPU boxes, so tests, the golden-fixture generator and bench.py all draw the same tensors from here.

Values depend only on (tensor name, shape, seed) — not on torch's RNG stream or parameter creation order — so the
reference model, the CPU oracle and the CUDA path can be filled identically on any machine.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

DIT_1_3B = dict(model_type="i2v", patch_size=(1, 2, 2), text_len=512, in_dim=36, dim=1536, ffn_dim=8960, freq_dim=256,
                text_dim=4096, out_dim=16, num_heads=12, num_layers=30, eps=1e-6)
# Small stand-in used by CPU-sized parity cases. dim/num_heads are fixed by the reference's audio adapter, which is
# hard-wired to 1536 channels (wan/models/wan_fantasy_transformer3d_1B.py:872) and by the 128-wide RoPE split.
DIT_TINY = dict(DIT_1_3B, ffn_dim=512, text_dim=128, text_len=24, num_layers=2)
# train_14B architecture (wan/models/wan_fantasy_transformer3d_14B.py:735, wan/configs/wan_i2v_14B.py:26-35): the adapter
# follows the DiT width (audio_proj_dim = dim, 8 heads) and has a two-stage audio projection 768 -> 2048 -> dim.
DIT_14B = dict(DIT_1_3B, dim=5120, ffn_dim=13824, num_heads=40, num_layers=40, variant="14B")
DIT_14B_TINY = dict(DIT_14B, dim=512, num_heads=4, ffn_dim=1024, text_dim=128, text_len=24, num_layers=2)


def _rng(name: str, seed: int) -> np.random.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return np.random.default_rng(int.from_bytes(h[:8], "little"))


def det_normal(name: str, shape, std: float = 1.0, mean: float = 0.0, seed: int = 0) -> torch.Tensor:
    a = _rng(name, seed).standard_normal(tuple(shape), dtype=np.float32) * np.float32(std) + np.float32(mean)
    return torch.from_numpy(a)


def fill_state_dict(shapes: dict, seed: int = 0) -> dict:
    """shapes: {reference parameter name: shape}. Linear/conv weights ~ N(0, 1/fan_in), biases ~ N(0, 0.02^2),
    norm scales ~ N(1, 0.1^2), modulation tables ~ N(0, 1/dim) (as the reference ctor, 1B.py:647). The reference
    zero-initialises cross_attn.{k,v}_vocal (1B.py:526-531); they are randomised here like every other Linear so
    the audio cross-attention is exercised."""
    sd = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        if name.endswith("modulation"):
            sd[name] = det_normal(name, shape, std=shape[-1] ** -0.5, seed=seed)
        elif name.endswith(".bias"):
            sd[name] = det_normal(name, shape, std=0.02, seed=seed)
        elif len(shape) == 1 or name.endswith(".gamma"):
            sd[name] = det_normal(name, shape, std=0.1, mean=1.0, seed=seed)
        else:
            fan_in = int(np.prod(shape[1:]))
            sd[name] = det_normal(name, shape, std=fan_in ** -0.5, seed=seed)
    return sd


def dit_param_shapes(cfg: dict) -> dict:
    """Parameter names/shapes of WanTransformer3DFantasyModel (1B.py:829-872; probe dump in SURVEY.md §8b)."""
    d, f, nl = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"]
    ps = cfg["patch_size"]
    s = {"patch_embedding.weight": (d, cfg["in_dim"], *ps), "patch_embedding.bias": (d,)}

    def lin(name, i, o, bias=True):
        s[name + ".weight"] = (o, i)
        if bias:
            s[name + ".bias"] = (o,)

    lin("text_embedding.0", cfg["text_dim"], d)
    lin("text_embedding.2", d, d)
    lin("time_embedding.0", cfg["freq_dim"], d)
    lin("time_embedding.2", d, d)
    lin("time_projection.1", d, 6 * d)
    for i in range(nl):
        p = f"blocks.{i}."
        s[p + "modulation"] = (1, 6, d)
        for n in ("q", "k", "v", "o"):
            lin(p + "self_attn." + n, d, d)
        s[p + "self_attn.norm_q.weight"] = (d,)
        s[p + "self_attn.norm_k.weight"] = (d,)
        s[p + "norm3.weight"] = (d,)
        s[p + "norm3.bias"] = (d,)
        for n in ("q", "k", "v", "o", "k_img", "v_img", "k_vocal", "v_vocal"):
            lin(p + "cross_attn." + n, d, d)
        for n in ("norm_q", "norm_k", "norm_k_img"):
            s[p + f"cross_attn.{n}.weight"] = (d,)
        lin(p + "ffn.0", d, f)
        lin(p + "ffn.2", f, d)
    s["head.modulation"] = (1, 2, d)
    lin("head.head", d, cfg["out_dim"] * int(np.prod(ps)))
    s["img_emb.proj.0.weight"] = (1280,)
    s["img_emb.proj.0.bias"] = (1280,)
    lin("img_emb.proj.1", 1280, 1280)
    lin("img_emb.proj.3", 1280, d)
    s["img_emb.proj.4.weight"] = (d,)
    s["img_emb.proj.4.bias"] = (d,)
    # audio adapter (vocal_projector_fantasy_1B.py:402-425): 768 -> 1536, 2 blocks, 8 heads, ffn 3072;
    # 14B (vocal_projector_fantasy_14B.py:384-425): 768 -> 2048 -> dim, blocks at the DiT width
    v = "vocal_projector."
    if cfg.get("variant") == "14B":
        a = d
        for n, (i, o) in (("1", (768, 2048)), ("2", (2048, a))):
            lin(f"{v}proj_model.proj_{n}", i, o, bias=False)
            s[f"{v}proj_model.norm_{n}.weight"] = (o,)
            s[f"{v}proj_model.norm_{n}.bias"] = (o,)
    else:
        a = 1536
        lin(v + "proj_model.proj", 768, a, bias=False)
        s[v + "proj_model.norm.weight"] = (a,)
        s[v + "proj_model.norm.bias"] = (a,)
    for i in range(2):
        p = f"{v}blocks.{i}."
        s[p + "modulation"] = (1, 6, a)
        s[p + "norm3.weight"] = (a,)
        s[p + "norm3.bias"] = (a,)
        lin(p + "cross_attn.q", a, a)
        lin(p + "cross_attn.k", d, a)
        lin(p + "cross_attn.v", d, a)
        lin(p + "cross_attn.o", a, a)
        s[p + "cross_attn.norm_q.weight"] = (a,)
        s[p + "cross_attn.norm_k.weight"] = (a,)
        lin(p + "ffn.0", a, 2 * a)
        lin(p + "ffn.2", 2 * a, a)
    s[v + "final_head.modulation"] = (1, 2, a)
    lin(v + "final_head.final_proj", a, a)
    return s


def dit_state_dict(cfg: dict, seed: int = 0) -> dict:
    return fill_state_dict(dit_param_shapes(cfg), seed)


def dit_inputs(cfg: dict, frames: int, height: int, width: int, batch: int = 3, audio_tokens: int | None = None,
               text_tokens: int = 16, t_value: float = 900.0, seed: int = 0) -> dict:
    """Inputs of one denoise evaluation as the pipeline builds them (pipe.py:726-750): the same latents for every CFG
    sample, y = [mask(4) | masked latents(16)], text = [neg, neg, prompt], audio = [0, a, a]."""
    F_lat, h, w = (frames - 1) // 4 + 1, height // 8, width // 8
    if audio_tokens is None:
        audio_tokens = frames * 2 - 1                       # wav2vec2 @16 kHz / 25 fps: 161 tokens for 81 frames
    lat = det_normal("latents", (1, 16, F_lat, h, w), seed=seed)
    msk = torch.zeros(1, 4, F_lat, h, w)
    msk[:, :, 0] = 1
    y = torch.cat([msk, det_normal("masked_latents", (1, 16, F_lat, h, w), seed=seed)], dim=1)
    neg = det_normal("text_neg", (text_tokens, cfg["text_dim"]), std=0.1, seed=seed)
    pos = det_normal("text_pos", (text_tokens + 3, cfg["text_dim"]), std=0.1, seed=seed)
    audio = det_normal("wav2vec", (1, audio_tokens, 768), seed=seed)
    if batch == 3:
        context = [neg, neg, pos]
        vocal = torch.cat([torch.zeros_like(audio), audio, audio])
    else:
        context = [pos] * batch
        vocal = audio.expand(batch, -1, -1).contiguous()
    return dict(x=lat.expand(batch, -1, -1, -1, -1).contiguous(), y=y.expand(batch, -1, -1, -1, -1).contiguous(),
                t=torch.full((batch,), t_value), context=context,
                clip_fea=det_normal("clip_fea", (1, 257, 1280), seed=seed).expand(batch, -1, -1).contiguous(),
                vocal_embeddings=vocal, seq_len=F_lat * (h // 2) * (w // 2), video_sample_n_frames=frames)


# ---------------------------------------------------------------------------------------------- Wan VAE (decode side)
def vae_decoder_layout(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_upsample=(True, True, False)):
    """Module list of Decoder3d.upsamples (wan/models/wan_vae.py:391-418): ('res', cin, cout) / ('up3d'|'up2d', c)."""
    dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
    mods = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(num_res_blocks + 1):
            mods.append(("res", cin, cout))
            cin = cout
        if i != len(dim_mult) - 1:
            mods.append(("up3d" if temperal_upsample[i] else "up2d", cout))
    return dims, mods


def vae_decoder_param_shapes(dim=96, z_dim=16, dim_mult=(1, 2, 4, 4)):
    """Names/shapes of the decode-side parameters of AutoencoderKLWan (wan/models/wan_vae.py:372-424, 516-518), checked
    against the real module by tools/gen_golden.py."""
    dims, mods = vae_decoder_layout(dim, dim_mult)
    s = {"model.conv2.weight": (z_dim, z_dim, 1, 1, 1), "model.conv2.bias": (z_dim,)}
    p = "model.decoder."
    s[p + "conv1.weight"] = (dims[0], z_dim, 3, 3, 3)
    s[p + "conv1.bias"] = (dims[0],)

    def res(pre, cin, cout):
        s[pre + "residual.0.gamma"] = (cin, 1, 1, 1)
        s[pre + "residual.2.weight"] = (cout, cin, 3, 3, 3)
        s[pre + "residual.2.bias"] = (cout,)
        s[pre + "residual.3.gamma"] = (cout, 1, 1, 1)
        s[pre + "residual.6.weight"] = (cout, cout, 3, 3, 3)
        s[pre + "residual.6.bias"] = (cout,)
        if cin != cout:
            s[pre + "shortcut.weight"] = (cout, cin, 1, 1, 1)
            s[pre + "shortcut.bias"] = (cout,)

    res(p + "middle.0.", dims[0], dims[0])
    s[p + "middle.1.norm.gamma"] = (dims[0], 1, 1)
    s[p + "middle.1.to_qkv.weight"] = (3 * dims[0], dims[0], 1, 1)
    s[p + "middle.1.to_qkv.bias"] = (3 * dims[0],)
    s[p + "middle.1.proj.weight"] = (dims[0], dims[0], 1, 1)
    s[p + "middle.1.proj.bias"] = (dims[0],)
    res(p + "middle.2.", dims[0], dims[0])
    c_last = dims[0]
    for i, m in enumerate(mods):
        q = f"{p}upsamples.{i}."
        if m[0] == "res":
            res(q, m[1], m[2])
            c_last = m[2]
        else:
            c = m[1]
            s[q + "resample.1.weight"] = (c // 2, c, 3, 3)
            s[q + "resample.1.bias"] = (c // 2,)
            if m[0] == "up3d":
                s[q + "time_conv.weight"] = (2 * c, c, 3, 1, 1)
                s[q + "time_conv.bias"] = (2 * c,)
            c_last = c // 2
    s[p + "head.0.gamma"] = (c_last, 1, 1, 1)
    s[p + "head.2.weight"] = (3, c_last, 3, 3, 3)
    s[p + "head.2.bias"] = (3,)
    return s


def vae_encoder_layout(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True)):
    """Module list of Encoder3d.downsamples (wan/models/wan_vae.py:294-310): ('res', cin, cout) / ('down3d'|'down2d', c)."""
    dims = [dim * u for u in [1] + list(dim_mult)]
    mods = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            mods.append(("res", cin, cout))
            cin = cout
        if i != len(dim_mult) - 1:
            mods.append(("down3d" if temperal_downsample[i] else "down2d", cout))
    return dims, mods


def vae_encoder_param_shapes(dim=96, z_dim=16, dim_mult=(1, 2, 4, 4)):
    """Names/shapes of the encode-side parameters of AutoencoderKLWan (wan/models/wan_vae.py:268-322, 515), checked
    against the real module by tools/gen_golden_vae.py."""
    dims, mods = vae_encoder_layout(dim, dim_mult)
    s = {"model.conv1.weight": (2 * z_dim, 2 * z_dim, 1, 1, 1), "model.conv1.bias": (2 * z_dim,)}
    p = "model.encoder."
    s[p + "conv1.weight"] = (dims[0], 3, 3, 3, 3)
    s[p + "conv1.bias"] = (dims[0],)

    def res(pre, cin, cout):
        s[pre + "residual.0.gamma"] = (cin, 1, 1, 1)
        s[pre + "residual.2.weight"] = (cout, cin, 3, 3, 3)
        s[pre + "residual.2.bias"] = (cout,)
        s[pre + "residual.3.gamma"] = (cout, 1, 1, 1)
        s[pre + "residual.6.weight"] = (cout, cout, 3, 3, 3)
        s[pre + "residual.6.bias"] = (cout,)
        if cin != cout:
            s[pre + "shortcut.weight"] = (cout, cin, 1, 1, 1)
            s[pre + "shortcut.bias"] = (cout,)

    for i, m in enumerate(mods):
        q = f"{p}downsamples.{i}."
        if m[0] == "res":
            res(q, m[1], m[2])
        else:
            c = m[1]
            s[q + "resample.1.weight"] = (c, c, 3, 3)
            s[q + "resample.1.bias"] = (c,)
            if m[0] == "down3d":
                s[q + "time_conv.weight"] = (c, c, 3, 1, 1)
                s[q + "time_conv.bias"] = (c,)
    c = dims[-1]
    res(p + "middle.0.", c, c)
    s[p + "middle.1.norm.gamma"] = (c, 1, 1)
    s[p + "middle.1.to_qkv.weight"] = (3 * c, c, 1, 1)
    s[p + "middle.1.to_qkv.bias"] = (3 * c,)
    s[p + "middle.1.proj.weight"] = (c, c, 1, 1)
    s[p + "middle.1.proj.bias"] = (c,)
    res(p + "middle.2.", c, c)
    s[p + "head.0.gamma"] = (c, 1, 1, 1)
    s[p + "head.2.weight"] = (2 * z_dim, c, 3, 3, 3)
    s[p + "head.2.bias"] = (2 * z_dim,)
    return s


def vae_state_dict(seed: int = 0, encoder: bool = False, **kw) -> dict:
    """Decoder-side state dict; with encoder=True the encoder + conv1 parameters too (values are seeded per name, so
    the decoder tensors are the same either way)."""
    shapes = dict(vae_decoder_param_shapes(**kw))
    if encoder:
        shapes.update(vae_encoder_param_shapes(**kw))
    return fill_state_dict(shapes, seed)
