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
