"""ORACLE (test infrastructure, not product code) — CPU restatement of the scheduler loop of
wan/pipeline/wan_inference_long_pipeline.py:704-791 and of diffusers==0.30.1's FlowMatchEulerDiscreteScheduler
(reference dependency, pyproject.toml:15; not vendored and not installed here).

PARITY UNPINNED for the scheduler arithmetic only: the reference holds no test or golden vector for it and diffusers
cannot be imported on this box; the restatement follows the published 0.30.1 source (tests/test_pipeline_cpu.py pins it
against a real diffusers wherever one is importable). The loop and the whole `__call__` chain ARE pinned:
tests/golden/pipeline_tiny.npz was written by the real reference pipeline class (tools/gen_golden_pipeline.py) and
`pipeline_call` below reproduces its latents to 4 bf16 ulp-flips in 5120 values. The loop is written with plain Python
lists / torch ops in the reference's statement order so that the product's restructured loop
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


def pipeline_call(dit_forward, vae_encode, vae_decode, cfg, *, tokenizer, text_encoder, clip_image_encoder, wav2vec_processor,
                  wav2vec, prompt, negative_prompt, height, width, clip_length, num_inference_steps, latents, vocal_input_values,
                  fps, sr, cond_file_path, overlap_window_length, text_guide_scale, audio_guide_scale, scheme="uniform",
                  max_sequence_length=512, return_latents=False):
    """Restatement of WanI2VTalkingInferenceLongPipeline.__call__ (pipe.py:540-806) for guidance_scale > 1 on the CPU
    oracles: `dit_forward(x, t, context, seq_len, clip_fea, y, vocal, video_sample_n_frames)`, `vae_encode(pixels)` ->
    [B, 32, T', h, w] (first 16 channels = mode), `vae_decode(latents)` -> [B, 3, T, H, W]. Pinned by
    tests/golden/pipeline_tiny.npz, which the REAL reference pipeline class wrote (tools/gen_golden_pipeline.py)."""
    from PIL import Image

    def t5(p):                                                                       # pipe.py:241-279
        tok = tokenizer([p], padding="max_length", max_length=max_sequence_length, truncation=True, add_special_tokens=True,
                        return_tensors="pt")
        lens = tok.attention_mask.gt(0).sum(dim=1).long()
        emb = text_encoder(tok.input_ids, attention_mask=tok.attention_mask)[0].float()
        return [u[:v] for u, v in zip(emb, lens)]
    pos, neg = t5(prompt), t5(negative_prompt or "")
    prompt_embeds = neg + neg + pos                                                  # pipe.py:636

    fpb = (clip_length - 1) // 4 + 1
    apf = int(sr / fps)
    max_audio_index = vocal_input_values.shape[0]
    latents_all = latents.clone()
    infer_length = latents_all.shape[2]

    img = Image.open(cond_file_path).convert("RGB").resize([width, height])          # pipe.py:661-673
    arr = (torch.from_numpy(np.array(img)).permute(2, 0, 1) / 255 - 0.5) * 2
    cond_image = arr.unsqueeze(1).unsqueeze(0)
    clip_context = torch.cat([clip_image_encoder([arr[:, None, :, :]])] * 3, dim=0)
    pixels = torch.cat([cond_image, torch.zeros(1, 3, clip_length - 1, height, width)], dim=2).float()
    masked = vae_encode(pixels)[:, :16]                                              # .mode(), pipe.py:402-403
    lh, lw = masked.shape[-2:]
    msk = torch.ones(1, clip_length, lh, lw)
    msk[:, 1:] = 0
    msk = torch.cat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], dim=1)
    msk = msk.view(1, msk.shape[1] // 4, 4, lh, lw).transpose(1, 2).float()
    y = torch.cat([torch.cat([msk] * 3), torch.cat([masked] * 3)], dim=1)
    seq_len = int(np.ceil((width // 8) * (height // 8) / 4 * fpb))                   # pipe.py:735

    def model_fn(lat, t, index_start, index_end, is_last):                           # pipe.py:714-750
        a0 = index_start * 4 * apf
        a1 = max_audio_index if is_last else a0 + (index_end - index_start) * 4 * apf
        sub = vocal_input_values[[ii % max_audio_index for ii in range(a0, a1)]]
        vals = wav2vec_processor(sub, sampling_rate=sr, return_tensors="pt").input_values
        feats = wav2vec(vals).last_hidden_state.float()
        vocal = torch.cat([torch.zeros_like(feats), feats, feats], dim=0)
        return dit_forward(torch.cat([lat] * 3), t.expand(3), prompt_embeds, seq_len, clip_context,
                           y[:, :, :lat.shape[2]], vocal, clip_length)

    # the reference rounds every window's latents to bf16 before writing them back, whatever the weight dtype (pipe.py:774, 779)
    lat = denoise_loop(model_fn, latents_all, num_inference_steps, clip_length, overlap_window_length, scheme=scheme,
                       audio_scale=audio_guide_scale, text_scale=text_guide_scale, dtype=torch.bfloat16)
    lat = lat.float()[:, :, :infer_length]
    if return_latents:
        return lat
    return (vae_decode(lat) / 2 + 0.5).clamp(0, 1)                                   # pipe.py:424-430
