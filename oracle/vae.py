"""ORACLE (test infrastructure, not product code) — CPU restatement of the Wan causal-3D VAE decode and encode.

Plain PyTorch fp32 on CPU over a reference-named state dict (keys as AutoencoderKLWan.state_dict(): "model.decoder...",
"model.conv2..."). Restates wan/models/wan_vae.py: CausalConv3d :20-39, RMS_norm :42-57, Resample(upsample2d/3d)
:69-143, ResidualBlock :189-223, AttentionBlock :226-265, Decoder3d :372-475, AutoencoderKLWan_.decode :549-574,
AutoencoderKLWan.decode :666-681. The per-conv causal feature cache is kept explicitly per conv index, exactly in the
order the reference's feat_idx counter walks the modules. Pinned by tests/golden/vae_tiny.npz (real reference output).
The encode side (Encoder3d :268-369, Resample(downsample2d/3d) :95-105, 145-162, AutoencoderKLWan_.encode :519-547,
AutoencoderKLWan.encode :649-664) is restated the same way and pinned by tests/golden/vae_enc_tiny.npz.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CACHE_T = 2
LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
               -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
              1.1253, 2.8251, 1.9160]


def decoder_layout(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_upsample=(True, True, False)):
    """Module list of Decoder3d.upsamples (vae.py:391-418): ('res', cin, cout) / ('up3d'|'up2d', c)."""
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


def causal_conv3d(x, w, b, cache):
    """vae.py:31-39 — left-pad time with the cached frames, zeros for whatever is still missing; H/W zero 'same' pad."""
    kt, kh, kw = w.shape[2:]
    pt = kt - 1
    if cache is not None and pt > 0:
        x = torch.cat([cache, x], dim=2)
        pt -= cache.shape[2]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, pt, 0))
    return F.conv3d(x, w, b)


def rms_norm(x, gamma):
    """vae.py:54-57 — F.normalize over channels * sqrt(C) * gamma."""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


class _Cache:
    def __init__(self, n):
        self.map, self.idx = [None] * n, 0


def _cached_conv(sd, name, x, fc):
    """The cache bookkeeping wrapped around every 3x3x3 CausalConv3d (vae.py:208-220, 428-440, 460-473)."""
    idx = fc.idx
    cache_x = x[:, :, -CACHE_T:].clone()
    if cache_x.shape[2] < 2 and fc.map[idx] is not None:
        cache_x = torch.cat([fc.map[idx][:, :, -1:], cache_x], dim=2)
    y = causal_conv3d(x, sd[name + ".weight"], sd[name + ".bias"], fc.map[idx])
    fc.map[idx] = cache_x
    fc.idx += 1
    return y


def residual_block(sd, pre, x, fc, has_shortcut):
    h = causal_conv3d(x, sd[pre + "shortcut.weight"], sd[pre + "shortcut.bias"], None) if has_shortcut else x
    x = F.silu(rms_norm(x, sd[pre + "residual.0.gamma"]))
    x = _cached_conv(sd, pre + "residual.2", x, fc)
    x = F.silu(rms_norm(x, sd[pre + "residual.3.gamma"]))
    x = _cached_conv(sd, pre + "residual.6", x, fc)
    return x + h


def attention_block(sd, pre, x):
    """vae.py:243-265 — one head of width C per frame."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = rms_norm(y, sd[pre + "norm.gamma"])
    qkv = F.conv2d(y, sd[pre + "to_qkv.weight"], sd[pre + "to_qkv.bias"])
    q, k, v = qkv.reshape(b * t, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
    a = torch.softmax(q @ k.transpose(-1, -2) / (c ** 0.5), dim=-1) @ v
    a = a.squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
    a = F.conv2d(a, sd[pre + "proj.weight"], sd[pre + "proj.bias"])
    return a.view(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x


def resample_up(sd, pre, x, fc, mode):
    """vae.py:106-143 — upsample3d: (3,1,1) causal time conv C->2C + frame interleave, skipped on the first chunk
    ('Rep'); then nearest-exact 2x + Conv2d 3x3 C->C/2 per frame."""
    b, c, t, h, w = x.shape
    if mode == "up3d":
        idx = fc.idx
        if fc.map[idx] is None:
            fc.map[idx] = "Rep"
            fc.idx += 1
        else:
            cache_x = x[:, :, -CACHE_T:].clone()
            prev = fc.map[idx]
            if cache_x.shape[2] < 2 and not isinstance(prev, str):
                cache_x = torch.cat([prev[:, :, -1:], cache_x], dim=2)
            if cache_x.shape[2] < 2 and isinstance(prev, str):
                cache_x = torch.cat([torch.zeros_like(cache_x), cache_x], dim=2)
            x = causal_conv3d(x, sd[pre + "time_conv.weight"], sd[pre + "time_conv.bias"],
                              None if isinstance(prev, str) else prev)
            fc.map[idx] = cache_x
            fc.idx += 1
            x = x.reshape(b, 2, c, t, h, w)
            x = torch.stack((x[:, 0], x[:, 1]), 3).reshape(b, c, t * 2, h, w)
    t = x.shape[2]
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact")
    y = F.conv2d(y, sd[pre + "resample.1.weight"], sd[pre + "resample.1.bias"], padding=1)
    return y.view(b, t, c // 2, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)


def decoder_chunk(sd, x, fc, mods, pre="model.decoder."):
    """Decoder3d.forward (vae.py:426-475) on one latent frame with the running feature cache."""
    fc.idx = 0
    x = _cached_conv(sd, pre + "conv1", x, fc)
    x = residual_block(sd, pre + "middle.0.", x, fc, False)
    x = attention_block(sd, pre + "middle.1.", x)
    x = residual_block(sd, pre + "middle.2.", x, fc, False)
    for i, m in enumerate(mods):
        p = f"{pre}upsamples.{i}."
        if m[0] == "res":
            x = residual_block(sd, p, x, fc, m[1] != m[2])
        else:
            x = resample_up(sd, p, x, fc, m[0])
    x = F.silu(rms_norm(x, sd[pre + "head.0.gamma"]))
    return _cached_conv(sd, pre + "head.2", x, fc)


def count_cached_convs(mods):
    return 1 + 4 + sum(2 if m[0] == "res" else (1 if m[0] == "up3d" else 0) for m in mods) + 1


def vae_decode(sd, z, dim=96, dim_mult=(1, 2, 4, 4), hooks=None):
    """AutoencoderKLWan.decode (vae.py:666-681) for z [B, 16, T, h, w] -> [B, 3, 1 + 4 (T-1), 8h, 8w] clamped [-1, 1]."""
    _, mods = decoder_layout(dim, dim_mult)
    mean = torch.tensor(LATENT_MEAN).view(1, -1, 1, 1, 1)
    inv_std = 1.0 / torch.tensor(LATENT_STD)
    outs = []
    for u in z:
        u = u.unsqueeze(0) / inv_std.view(1, -1, 1, 1, 1) + mean                      # vae.py:552-557
        x = causal_conv3d(u, sd["model.conv2.weight"], sd["model.conv2.bias"], None)
        fc = _Cache(count_cached_convs(mods))
        frames = []
        for i in range(x.shape[2]):
            o = decoder_chunk(sd, x[:, :, i:i + 1], fc, mods)
            if hooks is not None:
                hooks.setdefault("chunks", []).append(o)
            frames.append(o)
        outs.append(torch.cat(frames, dim=2).clamp_(-1, 1).squeeze(0))
    return torch.stack(outs)


# ------------------------------------------------------------------------------------------------------ encode
def encoder_layout(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True)):
    """Module list of Encoder3d.downsamples (vae.py:294-310): ('res', cin, cout) / ('down3d'|'down2d', c)."""
    dims = [dim * u for u in [1] + list(dim_mult)]
    mods = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            mods.append(("res", cin, cout))
            cin = cout
        if i != len(dim_mult) - 1:
            mods.append(("down3d" if temperal_downsample[i] else "down2d", cout))
    return dims, mods


def resample_down(sd, pre, x, fc, mode):
    """vae.py:145-162 — ZeroPad2d(right 1, bottom 1) + Conv2d 3x3 stride 2 per frame; downsample3d then applies a
    (3,1,1) stride-(2,1,1) conv over [last frame of the previous chunk, x] — except on the first chunk, whose frame
    passes through untouched and only seeds the cache."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.conv2d(F.pad(y, (0, 1, 0, 1)), sd[pre + "resample.1.weight"], sd[pre + "resample.1.bias"], stride=2)
    x = y.view(b, t, c, y.shape[-2], y.shape[-1]).permute(0, 2, 1, 3, 4)
    if mode == "down3d":
        idx = fc.idx
        if fc.map[idx] is None:
            fc.map[idx] = x.clone()
        else:
            cache_x = x[:, :, -1:].clone()
            x = F.conv3d(torch.cat([fc.map[idx][:, :, -1:], x], 2), sd[pre + "time_conv.weight"],
                         sd[pre + "time_conv.bias"], stride=(2, 1, 1))
            fc.map[idx] = cache_x
        fc.idx += 1
    return x


def encoder_chunk(sd, x, fc, mods, pre="model.encoder."):
    """Encoder3d.forward (vae.py:324-369) on one chunk of frames (1 for the first chunk, then 4)."""
    fc.idx = 0
    x = _cached_conv(sd, pre + "conv1", x, fc)
    for i, m in enumerate(mods):
        p = f"{pre}downsamples.{i}."
        if m[0] == "res":
            x = residual_block(sd, p, x, fc, m[1] != m[2])
        else:
            x = resample_down(sd, p, x, fc, m[0])
    x = residual_block(sd, pre + "middle.0.", x, fc, False)
    x = attention_block(sd, pre + "middle.1.", x)
    x = residual_block(sd, pre + "middle.2.", x, fc, False)
    x = F.silu(rms_norm(x, sd[pre + "head.0.gamma"]))
    return _cached_conv(sd, pre + "head.2", x, fc)


def count_cached_convs_enc(mods):
    return 1 + sum(2 if m[0] == "res" else (1 if m[0] == "down3d" else 0) for m in mods) + 4 + 1


def vae_encode(sd, x, dim=96, dim_mult=(1, 2, 4, 4), hooks=None):
    """AutoencoderKLWan._encode (vae.py:642-647) for x [B, 3, T, H, W] -> [B, 32, 1 + (T-1)//4, H/8, W/8]: channels
    [0,16) the normalised mean (what DiagonalGaussianDistribution.mode() returns), [16,32) the raw log-variance."""
    _, mods = encoder_layout(dim, dim_mult)
    zc = sd["model.conv1.weight"].shape[0] // 2
    mean = torch.tensor(LATENT_MEAN[:zc]).view(1, -1, 1, 1, 1)
    inv_std = (1.0 / torch.tensor(LATENT_STD[:zc])).view(1, -1, 1, 1, 1)
    outs = []
    for u in x:
        u = u.unsqueeze(0)
        fc = _Cache(count_cached_convs_enc(mods))
        chunks = []
        for i in range(1 + (u.shape[2] - 1) // 4):                                   # vae.py:523-538: 1, 4, 4, ...
            frames = u[:, :, :1] if i == 0 else u[:, :, 1 + 4 * (i - 1):1 + 4 * i]
            o = encoder_chunk(sd, frames, fc, mods)
            if hooks is not None:
                hooks.setdefault("chunks", []).append(o)
            chunks.append(o)
        out = torch.cat(chunks, dim=2)
        mu, log_var = causal_conv3d(out, sd["model.conv1.weight"], sd["model.conv1.bias"], None).chunk(2, dim=1)
        mu = (mu - mean) * inv_std                                                    # vae.py:540-542
        outs.append(torch.cat([mu, log_var], dim=1).squeeze(0))
    return torch.stack(outs)
