"""GPU tests (-m gpu) of the drop-in boundary at the level the reference's entry point uses it: the construction + call
sequence of inference.py:497-568 executed against the repo classes (encoders stubbed by the same deterministic functions
the golden generator fed the REAL reference pipeline, tools/pipeline_stubs.py), held against
  * tests/golden/pipeline_tiny.npz — written by the real WanI2VTalkingInferenceLongPipeline.__call__ (fp32 CPU), and
  * the oracle chain (oracle/pipeline.py) for a full 50-step sampler run (north_star: >= 35 dB after the full sampler)."""
import math

import numpy as np
import pytest
import torch

from stableavatar_b200 import synth

pytestmark = pytest.mark.gpu
CFG = synth.DIT_TINY
KEYS = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
        "num_heads", "num_layers")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def psnr(a, b):
    mse = ((torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()) ** 2).mean().item()
    return 10 * math.log10(1.0 / mse)


def build_pipeline(device="cuda"):
    """inference.py:470-524, statement for statement, with the repo's classes in place of the reference's."""
    from stableavatar_b200.pipeline import WanI2VTalkingInferenceLongPipeline
    from stableavatar_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    from stableavatar_b200.wan_vae import AutoencoderKLWan
    from tools import pipeline_stubs as S
    weight_dtype = torch.bfloat16
    tokenizer = S.Tokenizer()
    text_encoder = S.TextEncoder(CFG["text_dim"])
    vae = AutoencoderKLWan()
    vae.load_state_dict(synth.vae_state_dict(encoder=True), strict=True)
    wav2vec_processor, wav2vec = S.Wav2VecProcessor(), S.Wav2Vec()
    clip_image_encoder = S.ClipEncoder()
    transformer3d = WanTransformer3DFantasyModel(**{k: CFG[k] for k in KEYS})
    m, u = transformer3d.load_state_dict(synth.dit_state_dict(CFG), strict=False)          # inference.py:505
    assert not m and not u
    transformer3d = transformer3d.to(weight_dtype)                                          # from_pretrained(torch_dtype=...)
    scheduler = FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=5.0, use_dynamic_shifting=False)
    pipeline = WanI2VTalkingInferenceLongPipeline(
        tokenizer=tokenizer, text_encoder=text_encoder, vae=vae, transformer=transformer3d,
        clip_image_encoder=clip_image_encoder, scheduler=scheduler, wav2vec_processor=wav2vec_processor, wav2vec=wav2vec)
    pipeline.to(device=device)                                                              # inference.py:524
    return pipeline


def call(pipeline, c, steps=None, **extra):
    """inference.py:541-568 (+ `latents=` so the noise is the golden's)."""
    from tools import pipeline_stubs as S
    S.write_cond_image(c["cond_path"], c["height"], c["width"])
    generator = torch.Generator(device="cuda").manual_seed(43)
    with torch.no_grad():
        return pipeline(
            c["prompt"], num_frames=c["clip_length"], negative_prompt=c["negative_prompt"], height=c["height"], width=c["width"],
            guidance_scale=6.0, generator=generator, num_inference_steps=steps or c["steps"], video=None, mask_video=None,
            clip_image=None, text_guide_scale=c["text_scale"], audio_guide_scale=c["audio_scale"],
            vocal_input_values=c["audio"], motion_frame=25, fps=c["fps"], sr=c["sr"], cond_file_path=c["cond_path"], seed=43,
            overlap_window_length=c["overlap"], overlapping_weight_scheme=c["scheme"], clip_length=c["clip_length"],
            latents=c["latents"].clone(), **extra).videos


@pytest.fixture(scope="module")
def pipeline():
    return build_pipeline()


@pytest.mark.parametrize("name", ["windows3", "short_last"])
def test_inference_py_sequence_vs_real_reference_pipeline(pipeline, golden_dir, name):
    """Same construction, `.to(device=...)` and keyword call as inference.py — against what the REAL reference pipeline
    returned for the same inputs: latents of both scenarios within the bf16 bar, frames >= 35 dB."""
    from tools import pipeline_stubs as S
    gold = np.load(golden_dir / "pipeline_tiny.npz")
    c = S.case(name)
    assert pipeline.transformer.device.type == "cuda" and pipeline.vae.dtype == torch.float32
    pre = "" if name == "windows3" else name + "_"
    lat = call(pipeline, c, output_type="latent", return_dict=True)
    assert lat.dtype == torch.float32 and tuple(lat.shape) == gold[pre + "latents"].shape
    assert rel(lat, gold[pre + "latents"]) < 2e-2
    if name == "windows3":
        video = call(pipeline, c)
        assert isinstance(video, torch.Tensor) and tuple(video.shape) == (1, 3, 17, 64, 64)
        assert psnr(video, gold["video_f16"].astype(np.float32)) >= 35.0


def test_fifty_steps_frames_psnr_vs_oracle_chain(pipeline):
    """North-star bar: decoded frames >= 35 dB PSNR after the FULL sampler — 50 steps x 3 overlapping windows with 3-way
    CFG on the tiny model, bf16 product path against the fp32 CPU oracle chain (which reproduces the real reference
    pipeline, tests/test_pipeline_cpu.py)."""
    from oracle import dit as O, pipeline as OP, vae as OV
    from tools import pipeline_stubs as S
    c = S.case("windows3")
    video = call(pipeline, c, steps=50)
    sd = {k: v.bfloat16().float() for k, v in synth.dit_state_dict(CFG).items()}     # the weights the bf16 model holds
    sd_vae = synth.vae_state_dict(encoder=True)

    def dit_forward(x, t, context, seq_len, clip_fea, y, vocal, frames):
        return O.dit_forward(sd, CFG, x, t, context, seq_len, clip_fea, y, vocal, frames)
    with torch.no_grad():
        ref = OP.pipeline_call(dit_forward, lambda p: OV.vae_encode(sd_vae, p), lambda z: OV.vae_decode(sd_vae, z), CFG,
                               tokenizer=S.Tokenizer(), text_encoder=S.TextEncoder(CFG["text_dim"]), clip_image_encoder=S.ClipEncoder(),
                               wav2vec_processor=S.Wav2VecProcessor(), wav2vec=S.Wav2Vec(), prompt=c["prompt"],
                               negative_prompt=c["negative_prompt"], height=c["height"], width=c["width"],
                               clip_length=c["clip_length"], num_inference_steps=50, latents=c["latents"],
                               vocal_input_values=c["audio"], fps=c["fps"], sr=c["sr"], cond_file_path=c["cond_path"],
                               overlap_window_length=c["overlap"], text_guide_scale=c["text_scale"],
                               audio_guide_scale=c["audio_scale"], scheme=c["scheme"])
    assert tuple(video.shape) == tuple(ref.shape)
    p = psnr(video, ref)
    print(f"50-step PSNR vs oracle chain: {p:.1f} dB")
    assert p >= 35.0, p


def test_window_batching_and_graph_replay_do_not_change_results(pipeline):
    """The windows of a step batched into one forward + the fused blend kernel + CUDA-graph replay against the same
    pipeline run one window per forward, eagerly."""
    from tools import pipeline_stubs as S
    c = S.case("windows3")
    a = call(pipeline, c, output_type="latent", return_dict=True)
    n_graphs = len(pipeline._graphs)
    pipeline.max_windows_per_forward, pipeline.use_cuda_graphs = 1, False
    try:
        b = call(pipeline, c, output_type="latent", return_dict=True)
    finally:
        pipeline.max_windows_per_forward, pipeline.use_cuda_graphs = 4, True
    assert n_graphs >= 1 and rel(a, b) < 2e-3
    print("batched+graphed vs serial eager: bit-equal" if torch.equal(a, b) else f"rel {rel(a, b):.2e}")


def test_graph_cache_is_bounded_and_conditioning_is_refreshed(pipeline):
    """repeated calls must not capture a new graph per call (the key holds shapes only), and a call with a new
    prompt must see the new conditioning through the static buffers."""
    from tools import pipeline_stubs as S
    c = S.case("windows3")
    first = call(pipeline, c, output_type="latent", return_dict=True)
    n = len(pipeline._graphs)
    other = dict(c, prompt="someone else sings")
    changed = call(pipeline, other, output_type="latent", return_dict=True)
    again = call(pipeline, c, output_type="latent", return_dict=True)
    assert len(pipeline._graphs) == n <= pipeline.max_graphs
    assert torch.equal(first, again)
    assert rel(changed, first) > 1e-4
    fresh = call(build_pipeline(), other, output_type="latent", return_dict=True)
    assert torch.equal(changed, fresh)


@pytest.mark.parametrize("pred_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("scheme", ["uniform", "log"])
def test_window_blend_kernel_matches_the_reference_statements(pred_dtype, scheme):
    """sa_window_blend against pipe.py:756-779 executed with torch ops on the GPU, bit for bit: 5 windows of 21 frames with
    overlap 15 over 42 latent frames (the last window is short), bf16 and fp32 pred_latents."""
    from stableavatar_b200 import ops
    from stableavatar_b200.pipeline import overlap_weights, window_schedule
    N, fpb, overlap, C, h, w = 42, 21, 15, 16, 6, 10
    windows = window_schedule(N, fpb, overlap)
    g = torch.Generator(device="cuda").manual_seed(1)
    new = torch.randn(len(windows), C, fpb, h, w, generator=g, device="cuda").bfloat16()
    want = torch.zeros(1, C, N, h, w, device="cuda", dtype=pred_dtype)
    for k, (ws, we, prev_end) in enumerate(windows):
        latents = new[k:k + 1, :, :we - ws].clone()
        if ws != 0:
            ow = overlap_weights(overlap, scheme, "cuda", latents.dtype)
            s_idx = [ii % latents.shape[2] for ii in range(overlap)]
            e_idx = [ii % N for ii in range(prev_end - overlap, prev_end)]
            latents[:, :, s_idx] = (latents[:, :, s_idx] * ow + want[:, :, e_idx] * (1 - ow)).to(latents.dtype)
        latents = latents.to(torch.bfloat16)
        for iii in range(latents.shape[2]):
            want[:, :, (ws + iii) % N] = latents[:, :, iii]
    ow = overlap_weights(overlap, scheme, "cpu", torch.bfloat16).flatten()
    got = torch.zeros_like(want)
    ops.window_blend_(got, new, [(ws, we - ws, pe, ws != 0) for ws, we, pe in windows], overlap, ow.float().tolist(),
                      (1 - ow).float().tolist())
    assert torch.equal(got, want)


def test_cfg_euler_kernel_keeps_a_callers_fp32_sample():
    from stableavatar_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = torch.randn(3, 1000, generator=g, device="cuda").bfloat16()
    lat = torch.randn(1000, generator=g, device="cuda")                       # fp32, not bf16-representable
    u, d, c = pred
    noise = u + 5.0 * (d - u) + 3.0 * (c - d)
    want = (lat.float() + torch.tensor(-0.0123, dtype=torch.float32) * noise).to(torch.bfloat16)
    got = ops.cfg_euler_step(pred.contiguous(), lat.contiguous(), -0.0123, audio_scale=5.0, text_scale=3.0)
    assert got.dtype == torch.bfloat16 and torch.equal(got, want)
