"""ORACLE (test infrastructure, not product code) — CPU restatement of the scheduler loop of
wan/pipeline/wan_inference_long_pipeline.py:704-791 and of diffusers==0.30.1's FlowMatchEulerDiscreteScheduler
(reference dependency, pyproject.toml:15; not vendored and not installed here).

PARITY UNPINNED for the scheduler arithmetic: the reference holds no test or golden vector for it and diffusers
cannot be imported on this box; the restatement follows the published 0.30.1 source. The loop itself is written with
plain Python lists / torch ops exactly in the reference's statement order so that the product's restructured loop
(stableavatar_b200/pipeline.py) can be compared against it on any model callable.
"""
from __future__ import annotations

import numpy as np
import torch


def flow_match_sigmas(num_inference_steps, shift=5.0, num_train_timesteps=1000):
    """ctor + set_timesteps of FlowMatchEulerDiscreteScheduler (shift applied in the ctor AND in set_timesteps)."""
    t = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
    s = torch.from_numpy(t) / num_train_timesteps
    s = shift * s / (1 + (shift - 1) * s)
    smax, smin = s[0].item(), s[-1].item()
    ts = np.linspace(smax * num_train_timesteps, smin * num_train_timesteps, num_inference_steps)
    sig = ts / num_train_timesteps
    sig = shift * sig / (1 + (shift - 1) * sig)
    sig = torch.from_numpy(sig).to(torch.float32)
    return torch.cat([sig, torch.zeros(1)]), sig * num_train_timesteps


def euler_step(model_output, sample, sigma, sigma_next):
    """FlowMatchEulerDiscreteScheduler.step: fp32 update, cast back to the model output dtype."""
    return (sample.to(torch.float32) + (sigma_next - sigma) * model_output).to(model_output.dtype)


def cfg_combine(noise_pred, audio_scale, text_scale):
    """pipe.py:751-753."""
    u, d, c = noise_pred.chunk(3)
    return u + audio_scale * (d - u) + text_scale * (c - d)


def denoise_loop(model_fn, latents_all, num_inference_steps, clip_length, overlap, scheme="uniform", audio_scale=5.0,
                 text_scale=3.0, dtype=torch.float32):
    """pipe.py:704-791 with `model_fn(latents[1,C,f,h,w], t, index_start, index_end, is_last) -> noise_pred [3,...]`.
    The single-window case (infer_length == frames_per_batch), where the reference's while-loop never exits, is run
    once (see stableavatar_b200/pipeline.py)."""
    sigmas, timesteps = flow_match_sigmas(num_inference_steps)
    fpb = (clip_length - 1) // 4 + 1
    infer_length = latents_all.shape[2]
    for i, t in enumerate(timesteps):
        pred_latents = torch.zeros_like(latents_all)
        arrive_last = False
        index_start, index_end = 0, fpb
        index_previous_end = index_end
        while index_end <= infer_length:
            idx_list = [ii % latents_all.shape[2] for ii in range(index_start, index_end)]
            latents = latents_all[:, :, idx_list].clone()
            noise_pred = cfg_combine(model_fn(latents, t, index_start, index_end, index_end == infer_length), audio_scale, text_scale)
            latents = euler_step(noise_pred, latents, sigmas[i], sigmas[i + 1])
            if index_start != 0 and i != 0:
                w = torch.zeros(1, 1, overlap, 1, 1, dtype=latents.dtype)
                if scheme == "uniform":
                    for j in range(overlap):
                        w[:, :, j] = j / (overlap - 1)
                elif scheme == "log":
                    init = torch.log1p(torch.linspace(0, 1, overlap) * (torch.exp(torch.tensor(1.0)) - 1))
                    norm = (init - init.min()) / (init.max() - init.min())
                    for j in range(overlap):
                        w[:, :, j] = norm[j]
                s_idx = [ii % latents.shape[2] for ii in range(0, overlap)]
                e_idx = [ii % latents_all.shape[2] for ii in range(index_previous_end - overlap, index_previous_end)]
                latents[:, :, s_idx] = latents[:, :, s_idx] * w + pred_latents[:, :, e_idx] * (1 - w)
            latents = latents.to(dtype)
            for iii in range(latents.size(2)):
                pred_latents[:, :, (index_start + iii) % pred_latents.shape[2]] = latents[:, :, iii]
            if arrive_last or index_end == infer_length:
                break
            index_previous_end = index_end
            index_start = index_start + (fpb - overlap)
            if (index_start + fpb) < infer_length:
                index_end = index_start + fpb
            else:
                index_end = infer_length
                arrive_last = True
        latents_all = pred_latents
    return latents_all
