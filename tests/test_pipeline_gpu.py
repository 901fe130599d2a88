"""GPU parity (-m gpu) of the scheduler loop: stableavatar_b200.pipeline.denoise (CUDA DiT + fused CFG/Euler kernel,
sliding windows with overlap blending) against the oracle restatement of the reference loop driving the CPU oracle DiT."""
import pytest
import torch

from stableavatar_b200 import synth

pytestmark = pytest.mark.gpu
CFG = synth.DIT_TINY


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_two_steps_three_windows_vs_oracle_loop():
    from oracle import dit as O, pipeline as OP
    from stableavatar_b200.pipeline import WanI2VTalkingInferenceLongPipeline
    from stableavatar_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    sd = {k: v.bfloat16() for k, v in synth.dit_state_dict(CFG).items()}
    model = WanTransformer3DFantasyModel(**{k: CFG[k] for k in keys})
    model.load_state_dict(sd, strict=True)
    model = model.to("cuda", torch.bfloat16)
    pipe = WanI2VTalkingInferenceLongPipeline(transformer=model, scheduler=FlowMatchEulerDiscreteScheduler(1000, 5.0))

    clip_length, overlap, steps = 9, 2, 2                       # 3 latent frames per window, windows [0,3) [1,4) [2,5)
    inp = synth.dit_inputs(CFG, frames=clip_length, height=64, width=96)
    r = lambda t: t.bfloat16().float()  # noqa: E731
    lat_all = r(synth.det_normal("latents_all", (1, 16, 5, 8, 12)))      # 5 latent frames -> windows [0,3) [2,5)
    audio = {(ws, ws + 3): r(synth.det_normal(f"a{ws}", (1, 17, 768))) for ws in (0, 1, 2)}
    dev, bf = "cuda", torch.bfloat16

    out = pipe.denoise(lat_all.to(dev, bf), [c.to(dev, bf) for c in inp["context"]], inp["clip_fea"].to(dev, bf),
                       inp["y"].to(dev, bf), lambda ws, we, last: audio[(ws, we)], num_inference_steps=steps,
                       clip_length=clip_length, text_guide_scale=3.0, audio_guide_scale=5.0,
                       overlap_window_length=overlap, seq_len=inp["seq_len"])
    torch.cuda.synchronize()

    sdf = {k: v.float() for k, v in sd.items()}

    def model_fn(latents, t, ws, we, last):
        a = audio[(ws, we)]
        with torch.no_grad():
            return O.dit_forward(sdf, CFG, latents.expand(3, -1, -1, -1, -1), t.expand(3), [r(c) for c in inp["context"]],
                                 inp["seq_len"], r(inp["clip_fea"]), r(inp["y"])[:, :, :latents.shape[2]],
                                 torch.cat([torch.zeros_like(a), a, a]), clip_length)
    ref = OP.denoise_loop(model_fn, lat_all.clone(), steps, clip_length, overlap)
    assert out.shape == ref.shape
    assert rel(out, ref) < 2e-2


def test_cfg_euler_kernel_matches_bf16_torch_ops():
    """sa_cfg_euler_step reproduces the reference's op-by-op bf16 rounding (pipe.py:751-754) bit for bit."""
    from stableavatar_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = torch.randn(3, 1000, generator=g, device="cuda").bfloat16()
    lat = torch.randn(1000, generator=g, device="cuda").bfloat16()
    u, d, c = pred
    noise = u + 5.0 * (d - u) + 3.0 * (c - d)
    want = (lat.float() + torch.tensor(-0.0123, dtype=torch.float32) * noise).to(torch.bfloat16)
    got = ops.cfg_euler_step(pred.contiguous(), lat.contiguous(), -0.0123, audio_scale=5.0, text_scale=3.0)
    assert torch.equal(got, want)


def test_call_end_to_end_frames_psnr_vs_oracle():
    """The reference entry point `pipe(...)` (pipe.py:540-806) end to end on the B200 path — VAE encode of the conditioning
    clip, mask / y assembly, 3 denoise steps with 3-way CFG, VAE decode, `/ 2 + 0.5` — against the same chain built from
    the CPU oracles. North-star bar: decoded frames >= 35 dB PSNR after the full sampler."""
    import math
    from oracle import dit as O, pipeline as OP, vae as OV
    from stableavatar_b200.pipeline import WanI2VTalkingInferenceLongPipeline
    from stableavatar_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    from stableavatar_b200.wan_vae import AutoencoderKLWan
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    sd = {k: v.bfloat16() for k, v in synth.dit_state_dict(CFG).items()}
    model = WanTransformer3DFantasyModel(**{k: CFG[k] for k in keys})
    model.load_state_dict(sd, strict=True)
    model = model.to("cuda", torch.bfloat16)
    sd_vae = synth.vae_state_dict(encoder=True)
    vae = AutoencoderKLWan()
    vae.load_state_dict(sd_vae, strict=True)
    vae = vae.to("cuda")
    pipe = WanI2VTalkingInferenceLongPipeline(vae=vae, transformer=model, scheduler=FlowMatchEulerDiscreteScheduler(1000, 5.0))

    H = W = 64
    frames, steps = 9, 3
    r = lambda t: t.bfloat16().float()  # noqa: E731
    pos, neg = r(synth.det_normal("e2e_pos", (12, CFG["text_dim"]), std=0.1)), r(synth.det_normal("e2e_neg", (9, CFG["text_dim"]), std=0.1))
    clip = r(synth.det_normal("e2e_clip", (1, 257, 1280)))
    cond = synth.det_normal("e2e_img", (1, 3, 1, H, W)).clamp_(-1, 1)
    lat0 = r(synth.det_normal("e2e_lat", (1, 16, 3, H // 8, W // 8)))
    audio = r(synth.det_normal("e2e_audio", (1, 2 * frames - 1, 768)))
    out = pipe(height=H, width=W, num_frames=frames, clip_length=frames, num_inference_steps=steps, guidance_scale=6.0,
               text_guide_scale=3.0, audio_guide_scale=5.0, latents=lat0, prompt_embeds=[pos], negative_prompt_embeds=[neg],
               clip_context=clip, cond_image=cond, vocal_input_values=torch.zeros(frames * 640), sr=16000, fps=25,
               vocal_embeddings_fn=lambda ws, we, last: audio, overlap_window_length=2)
    video = out.videos
    assert tuple(video.shape) == (1, 3, frames, H, W) and float(video.min()) >= 0.0 and float(video.max()) <= 1.0

    with torch.no_grad():
        pixels = torch.cat([cond, torch.zeros(1, 3, frames - 1, H, W)], dim=2)
        masked = OV.vae_encode(sd_vae, pixels)[:, :16]
        lh, lw = masked.shape[-2:]
        msk = torch.ones(1, frames, lh, lw)
        msk[:, 1:] = 0
        msk = torch.cat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], dim=1)
        msk = msk.view(1, msk.shape[1] // 4, 4, lh, lw).transpose(1, 2)
        y = r(torch.cat([torch.cat([msk] * 3), torch.cat([masked] * 3)], dim=1))
        sdf = {k: v.float() for k, v in sd.items()}
        seq_len = math.ceil((W // 8) * (H // 8) / 4 * 3)

        def model_fn(latents, t, ws, we, last):
            return O.dit_forward(sdf, CFG, latents.expand(3, -1, -1, -1, -1), t.expand(3), [neg, neg, pos], seq_len,
                                 clip.expand(3, -1, -1), y[:, :, :latents.shape[2]], torch.cat([torch.zeros_like(audio), audio, audio]),
                                 frames)
        lat = OP.denoise_loop(model_fn, lat0.clone(), steps, frames, 2)
        ref = (OV.vae_decode(sd_vae, lat) / 2 + 0.5).clamp(0, 1)
    mse = ((video.double() - ref.double()) ** 2).mean().item()
    psnr = 10 * math.log10(1.0 / mse)
    assert psnr >= 35.0, psnr
