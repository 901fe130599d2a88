"""Golden fixture of the Wan VAE decode from the REAL reference module (wan/models/wan_vae.py), see gen_golden.py."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.refstub import import_reference  # noqa: E402
from stableavatar_b200 import synth  # noqa: E402


def gen_vae():
    _, _, vae = import_reference()
    m = vae.AutoencoderKLWan().eval()
    ref_sd = m.state_dict()
    sd = synth.vae_state_dict()
    dec_keys = {k: tuple(v.shape) for k, v in ref_sd.items() if k.startswith("model.decoder.") or k.startswith("model.conv2.")}
    mine = {k: tuple(v) for k, v in synth.vae_decoder_param_shapes().items()}
    assert dec_keys == mine, (set(dec_keys) ^ set(mine), [(k, dec_keys[k], mine[k]) for k in dec_keys if k in mine and dec_keys[k] != mine[k]])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("model.encoder.") or k.startswith("model.conv1.") for k in missing)
    res = {}
    z = synth.det_normal("vae_z", (1, 16, 3, 6, 8))
    with torch.no_grad():
        out = m.decode(z).sample
    res["z3_out"] = out.numpy()
    z1 = synth.det_normal("vae_z1", (2, 16, 1, 4, 6))                # single latent frame (first-chunk 'Rep' path), batch 2
    with torch.no_grad():
        res["z1_out"] = m.decode(z1).sample.numpy()
    res["config"] = np.array([m.config.latent_channels, m.config.temporal_compression_ratio, m.config.spacial_compression_ratio])
    np.savez_compressed(ROOT / "tests" / "golden" / "vae_tiny.npz", **res)
    print("wrote vae_tiny.npz", {k: v.shape for k, v in res.items()})


def gen_vae_encode():
    """Encode fixture: the REAL AutoencoderKLWan.encode on a 9-frame clip (chunks 1, 4, 4 -> 3 latent frames), on a
    single frame (first-chunk path only), and on a 6-frame clip (one trailing frame dropped by the 1+4k chunking)."""
    _, _, vae = import_reference()
    m = vae.AutoencoderKLWan().eval()
    ref_sd = m.state_dict()
    sd = synth.vae_state_dict(encoder=True)
    enc_keys = {k: tuple(v.shape) for k, v in ref_sd.items() if k.startswith("model.encoder.") or k.startswith("model.conv1.")}
    mine = {k: tuple(v) for k, v in synth.vae_encoder_param_shapes().items()}
    assert enc_keys == mine, (set(enc_keys) ^ set(mine), [(k, enc_keys[k], mine[k]) for k in enc_keys if k in mine and enc_keys[k] != mine[k]])
    missing, unexpected = m.load_state_dict(sd, strict=True)
    res = {}
    with torch.no_grad():
        x9 = synth.det_normal("vae_x9", (1, 3, 9, 32, 48)).clamp_(-1, 1)
        d = m.encode(x9).latent_dist
        res["x9_params"] = torch.cat([d.mean, d.logvar], 1).numpy()
        res["x9_mode"] = d.mode().numpy()
        x1 = synth.det_normal("vae_x1", (2, 3, 1, 16, 32)).clamp_(-1, 1)
        res["x1_mode"] = m.encode(x1).latent_dist.mode().numpy()
        x6 = synth.det_normal("vae_x6", (1, 3, 6, 16, 16)).clamp_(-1, 1)
        res["x6_mode"] = m.encode(x6).latent_dist.mode().numpy()
    np.savez_compressed(ROOT / "tests" / "golden" / "vae_enc_tiny.npz", **res)
    print("wrote vae_enc_tiny.npz", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    if "--encode-only" not in sys.argv:
        gen_vae()
    gen_vae_encode()


def gen_vae_fullres():
    """SURVEY.md §8c pin #4: the REAL AutoencoderKLWan.decode at the benchmark's latent grid — z [1,16,3,60,104] ->
    [1,3,9,480,832] (about 100 s on 8 host threads) — kept as a strided subsample (every 8th row / column: 60 x 104 pixels per
    frame, all 9 frames and 3 channels, fp16) plus per-frame means so that the fixture stays small (< 400 KB)."""
    _, _, vae = import_reference()
    m = vae.AutoencoderKLWan().eval()
    missing, unexpected = m.load_state_dict(synth.vae_state_dict(), strict=False)
    assert not unexpected
    z = synth.det_normal("vae_z_full", (1, 16, 3, 60, 104))
    with torch.no_grad():
        out = m.decode(z).sample
    assert tuple(out.shape) == (1, 3, 9, 480, 832)
    res = {"sub8": out[..., 3::8, 5::8].numpy().astype(np.float16), "frame_mean": out.mean(dim=(-1, -2)).numpy(),
           "frame_sq": (out.double() ** 2).mean(dim=(-1, -2)).numpy()}
    np.savez_compressed(ROOT / "tests" / "golden" / "vae_fullres_sub.npz", **res)
    print("wrote vae_fullres_sub.npz", {k: v.shape for k, v in res.items()})
