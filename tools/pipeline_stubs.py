"""Deterministic stand-ins for the context producers around the pipeline (T5 tokenizer / encoder, CLIP image encoder,
Wav2Vec2 processor / model) and one small end-to-end case. They are pure functions of their inputs, so the golden
generator (tools/gen_golden_pipeline.py, real reference pipeline) and the tests (oracle chain, product pipeline) feed
both sides exactly the same conditioning. None of this is part of the hot path."""
from __future__ import annotations

import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from stableavatar_b200 import synth


class Tokenizer:
    """Characters -> ids in [1, 97]; 0 is padding. Mirrors the keyword surface the pipeline uses (pipe.py:249-260)."""

    def __call__(self, prompt, padding="max_length", max_length=None, truncation=False, add_special_tokens=True,
                 return_tensors="pt"):
        prompt = [prompt] if isinstance(prompt, str) else list(prompt)
        rows = [[ord(ch) % 97 + 1 for ch in p] for p in prompt]
        width = max_length if (padding == "max_length" and max_length) else max(len(r) for r in rows)
        ids = torch.zeros(len(rows), width, dtype=torch.long)
        for i, r in enumerate(rows):
            r = r[:width]
            ids[i, :len(r)] = torch.tensor(r, dtype=torch.long)
        return SimpleNamespace(input_ids=ids, attention_mask=(ids > 0).long())

    def batch_decode(self, ids):
        return ["" for _ in ids]


class TextEncoder:
    dtype = torch.float32

    def __init__(self, text_dim):
        self.table = synth.det_normal("stub_t5_table", (98, text_dim), std=0.1)

    def __call__(self, ids, attention_mask=None):
        pos = torch.arange(ids.shape[1], dtype=torch.float32).view(1, -1, 1)
        return (self.table[ids.cpu()] + 0.01 * torch.sin(0.3 * pos),)


class ClipEncoder:
    def __init__(self):
        self.base = synth.det_normal("stub_clip", (1, 257, 1280))

    def __call__(self, images):
        img = images[0].float().cpu()                        # [3, 1, H, W] in [-1, 1]
        m = img.mean(dim=(1, 2, 3))                          # per-channel mean
        return self.base * (1.0 + 0.5 * m.mean()) + 0.1 * m[0]


class Wav2VecProcessor:
    def __call__(self, values, sampling_rate=16000, return_tensors="pt"):
        x = torch.as_tensor(values, dtype=torch.float32).reshape(1, -1)
        x = (x - x.mean()) / torch.sqrt(x.var(unbiased=False) + 1e-7)
        return SimpleNamespace(input_values=x)


class Wav2Vec:
    """T = N // 320 - 1 tokens (the stride / receptive field of wav2vec2-base), 768 channels."""

    def __call__(self, input_values):
        x = input_values.float().cpu().reshape(-1)
        T = x.numel() // 320 - 1
        win = torch.stack([x[t * 320:(t + 2) * 320] for t in range(T)])           # [T, 640]
        m, s = win.mean(dim=1, keepdim=True), win.std(dim=1, keepdim=True)
        c = torch.arange(1, 769, dtype=torch.float32).view(1, -1)
        feat = torch.sin(0.37 * c * m) + 0.5 * torch.cos(0.011 * c) * s + 0.05 * torch.sin(0.9 * c)
        return SimpleNamespace(last_hidden_state=feat.unsqueeze(0))


def write_cond_image(path, height, width):
    from PIL import Image
    yy, xx = np.mgrid[0:height, 0:width]
    img = np.stack([(yy * 3 + xx) % 256, (xx * 5 + 40) % 256, ((yy + xx) * 2) % 256], axis=-1).astype(np.uint8)
    Image.fromarray(img).save(path)


def case(name="windows3"):
    """17 frames @ 64x64 (5 latent frames), 2 steps so that the overlap blend (step > 0, window > 0) is exercised.
    "windows3": 9-frame windows (3 latent frames), overlap 2, uniform blend -> windows [0,3) [1,4) [2,5).
    "short_last": 13-frame windows (4 latent frames), overlap 2, log blend -> windows [0,4) [2,5): the last window holds
    only 3 latent frames, so the DiT sees zero-padded but live tokens and the audio runs to the end."""
    fps, sr, frames = 25, 16000, 17
    n = frames * (sr // fps)
    t = torch.arange(n, dtype=torch.float32) / sr
    audio = 0.4 * torch.sin(2 * np.pi * 220 * t) + 0.2 * torch.sin(2 * np.pi * 523 * t + 1.0) * torch.cos(2 * np.pi * 3 * t)
    lat = synth.det_normal("pipe_case_latents", (1, 16, 5, 8, 8)).bfloat16().float()
    c = dict(height=64, width=64, clip_length=9, steps=2, fps=fps, sr=sr, audio=audio, latents=lat, overlap=2, scheme="uniform",
             prompt="a person is talking", negative_prompt="blurry", text_scale=3.0,
             audio_scale=5.0, cond_path=str(Path(tempfile.gettempdir()) / "sa_b200_cond_image.png"))
    if name == "short_last":
        c.update(clip_length=13, scheme="log", text_scale=2.0, audio_scale=4.0)
    return c
