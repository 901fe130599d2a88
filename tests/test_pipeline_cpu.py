"""CPU tests of the host-side loop logic (no GPU): window schedule, overlap weights and scheduler tables of
stableavatar_b200/pipeline.py + scheduler.py against the oracle restatement of the reference loop."""
import numpy as np
import pytest
import torch

from oracle import pipeline as OP
from stableavatar_b200.pipeline import overlap_weights, window_schedule
from stableavatar_b200.scheduler import FlowMatchEulerDiscreteScheduler


def oracle_windows(infer_length, fpb, overlap):
    seen = []
    lat = torch.zeros(1, 1, infer_length, 1, 1)

    def fn(latents, t, ws, we, is_last):
        seen.append((ws, we, is_last))
        return torch.zeros(3, *latents.shape[1:])
    OP.denoise_loop(fn, lat, 1, (fpb - 1) * 4 + 1, overlap)
    return seen


@pytest.mark.parametrize("infer_length,fpb,overlap", [(42, 21, 15), (21, 21, 5), (30, 21, 5), (64, 21, 10), (43, 21, 15), (22, 21, 3)])
def test_window_schedule_matches_reference_loop(infer_length, fpb, overlap):
    want = oracle_windows(infer_length, fpb, overlap)
    got = window_schedule(infer_length, fpb, overlap)
    assert [(a, b) for a, b, _ in got] == [(a, b) for a, b, _ in want]
    assert got[-1][1] == infer_length


def test_appendix_c_example():
    # 42 latent frames, 21 per window, overlap 15
    assert [(a, b) for a, b, _ in window_schedule(42, 21, 15)] == [(0, 21), (6, 27), (12, 33), (18, 39), (24, 42)]


def test_no_window_when_clip_shorter_than_a_window():
    assert window_schedule(10, 21, 5) == []


@pytest.mark.parametrize("n", [10, 50])
def test_sigma_table(n):
    s = FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=5.0)
    s.set_timesteps(n, device="cpu", mu=1)
    sig, ts = OP.flow_match_sigmas(n)
    assert torch.equal(s.sigmas, sig) and torch.equal(s.timesteps.cpu(), ts)
    assert s.sigmas[-1] == 0 and s.sigmas[0] > 0.99 and (s.sigmas[:-1] > s.sigmas[1:]).all()
    for i in (0, n // 2, n - 1):
        assert s.index_for_timestep(s.timesteps[i]) == i
        assert s.dsigma_at(i) == float(sig[i + 1] - sig[i])


def test_overlap_weights():
    w = overlap_weights(5, "uniform", "cpu", torch.float32).flatten()
    assert torch.allclose(w, torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0]))
    w = overlap_weights(5, "log", "cpu", torch.float32).flatten()
    assert w[0] == 0 and abs(w[-1].item() - 1) < 1e-6 and (w[1:] > w[:-1]).all()


def test_vae_pipeline_partition():
    """Contiguous min-max partition used by the pipeline-parallel VAE decode."""
    from stableavatar_b200.wan_vae import AutoencoderKLWan
    part = AutoencoderKLWan.partition_units
    costs = [1, 1, 1, 1, 8, 1, 1, 1, 1]
    for stages in (1, 2, 3, 4, 9, 12):
        r = part(costs, stages)
        assert len(r) == stages and r[0][0] == 0
        used = [x for x in r if x[1] > x[0]]
        assert used[-1][1] == len(costs) and all(a[1] == b[0] for a, b in zip(used, used[1:]))
    assert max(sum(costs[a:b]) for a, b in part(costs, 3)) == 8
    assert max(sum(costs[a:b]) for a, b in part(costs, 2)) == 12
    assert part([5.0], 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]


def test_frames_kwarg_only_where_the_forward_takes_it():
    """The 14B class's forward has no video_sample_n_frames (14B.py:922-933): the pipeline must not pass it, and must
    refuse window lengths other than the hard-wired 81 frames."""
    import pytest
    from stableavatar_b200.pipeline import _frames_kwarg
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasy14BModel, WanTransformer3DFantasyModel

    class Obj:
        pass
    one, big = Obj(), Obj()
    one.forward = WanTransformer3DFantasyModel.forward.__get__(one)
    big.forward = WanTransformer3DFantasy14BModel.forward.__get__(big)
    assert _frames_kwarg(one, 33) == {"video_sample_n_frames": 33}
    assert _frames_kwarg(big, 81) == {}
    with pytest.raises(ValueError, match="81-frame"):
        _frames_kwarg(big, 33)
    assert _frames_kwarg(lambda **kw: None, 9) == {"video_sample_n_frames": 9}


@pytest.mark.parametrize("n", [10, 50])
def test_scheduler_tables_against_real_diffusers_when_installed(n):
    """the scheduler restatement is 'parity unpinned' because diffusers (reference dependency, pinned
    0.30.1) is not in this image. Wherever a real diffusers is importable, pin sigma / timestep tables and one step."""
    import sys
    mod = sys.modules.get("diffusers")
    if mod is not None and getattr(mod, "_sa_stub", False):
        pytest.skip("only the oracle's import stub of diffusers is present")
    diffusers = pytest.importorskip("diffusers")
    from oracle import pipeline as OP
    from stableavatar_b200.scheduler import FlowMatchEulerDiscreteScheduler
    ref = diffusers.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=5.0)
    ref.set_timesteps(n)
    mine = FlowMatchEulerDiscreteScheduler(1000, 5.0)
    mine.set_timesteps(n)
    sig, ts = OP.flow_match_sigmas(n)
    assert torch.allclose(ref.sigmas.float(), sig, atol=1e-6) and torch.allclose(ref.timesteps.float(), ts, atol=1e-3)
    assert torch.allclose(mine.sigmas.float().cpu(), sig, atol=1e-6)
    x, v = torch.randn(1, 4, 2, 3, 3), torch.randn(1, 4, 2, 3, 3)
    want = ref.step(v, ref.timesteps[0], x, return_dict=False)[0]
    assert torch.allclose(OP.euler_step(v, x, sig[0], sig[1]), want, atol=1e-6)


@pytest.mark.parametrize("name", ["windows3", "short_last"])
def test_oracle_pipeline_chain_vs_real_reference_pipeline(golden_dir, name):
    """tests/golden/pipeline_tiny.npz was written by the REAL WanI2VTalkingInferenceLongPipeline.__call__ (fp32, CPU; real
    DiT and VAE classes, stubbed context producers — tools/gen_golden_pipeline.py): 17 frames, 2 steps, 3-way CFG, VAE
    encode of the conditioning clip and VAE decode. "windows3": three overlapping 9-frame windows, uniform blend;
    "short_last": 13-frame windows whose last one holds 3 of 4 latent frames (live zero-pad tokens, audio to the end), log
    blend. The oracle chain must reproduce both."""
    import math
    from oracle import dit as O, vae as OV
    from stableavatar_b200 import synth
    from tools import pipeline_stubs as S
    gold = np.load(golden_dir / "pipeline_tiny.npz")
    cfg = synth.DIT_TINY
    sd, sd_vae = synth.dit_state_dict(cfg), synth.vae_state_dict(encoder=True)
    c = S.case(name)
    S.write_cond_image(c["cond_path"], c["height"], c["width"])

    def dit_forward(x, t, context, seq_len, clip_fea, y, vocal, frames):
        return O.dit_forward(sd, cfg, x, t, context, seq_len, clip_fea, y, vocal, frames)
    kw = dict(tokenizer=S.Tokenizer(), text_encoder=S.TextEncoder(cfg["text_dim"]), clip_image_encoder=S.ClipEncoder(),
              wav2vec_processor=S.Wav2VecProcessor(), wav2vec=S.Wav2Vec(), prompt=c["prompt"],
              negative_prompt=c["negative_prompt"], height=c["height"], width=c["width"], clip_length=c["clip_length"],
              num_inference_steps=c["steps"], latents=c["latents"], vocal_input_values=c["audio"], fps=c["fps"], sr=c["sr"],
              cond_file_path=c["cond_path"], overlap_window_length=c["overlap"], text_guide_scale=c["text_scale"],
              audio_guide_scale=c["audio_scale"], scheme=c["scheme"])
    with torch.no_grad():
        lat = OP.pipeline_call(dit_forward, lambda p: OV.vae_encode(sd_vae, p), lambda z: OV.vae_decode(sd_vae, z), cfg,
                               return_latents=True, **kw)
    pre = "" if name == "windows3" else name + "_"
    ref_lat = torch.from_numpy(gold[pre + "latents"])
    assert lat.shape == ref_lat.shape == (1, 16, 5, 8, 8)
    err = ((lat.double() - ref_lat.double()).norm() / ref_lat.double().norm()).item()
    assert err < 2e-3, err                      # bf16 write-back of every window: an occasional 1-ulp flip is all that may differ
    if name == "windows3":
        with torch.no_grad():
            video = (OV.vae_decode(sd_vae, lat) / 2 + 0.5).clamp(0, 1)
        ref_video = torch.from_numpy(gold["video_f16"].astype(np.float32))
        mse = ((video.double() - ref_video.double()) ** 2).mean().item()
        assert 10 * math.log10(1.0 / mse) > 45.0


# ---------------------------------------------------------------------------------------------- product host logic vs goldens
@pytest.mark.parametrize("T,nf", [(9, 5), (17, 9), (134, 81), (161, 81), (173, 81), (161, 69)])
def test_product_window_gather_table_vs_reference_golden(golden_dir, T, nf):
    """The PRODUCT's audio-window index table (stableavatar_b200/vocal_projector.py: split_audio_sequence +
    window_gather_table, what the gather kernel consumes) against the tables the REAL reference produced
    (vocal_projector_fantasy.py:39-131, tests/golden/dit_tiny.npz written by tools/gen_golden.py)."""
    from stableavatar_b200.vocal_projector import split_audio_sequence, window_gather_table
    gold = np.load(golden_dir / "dit_tiny.npz")
    ranges = split_audio_sequence(T, num_frames=nf)
    assert np.array_equal(np.array(ranges), gold[f"win_{T}_{nf}_ranges"])
    table, lens = window_gather_table(T, ranges, expand_length=4)
    # golden: 1-based source index of every slot, 0 = zero padding appended at the end of the window
    assert np.array_equal(np.array(table, dtype=np.int64) + 1, gold[f"win_{T}_{nf}_gather"])
    assert np.array_equal(np.array(lens), gold[f"win_{T}_{nf}_lens"])


def _bare_pipeline():
    from stableavatar_b200.pipeline import WanI2VTalkingInferenceLongPipeline
    return WanI2VTalkingInferenceLongPipeline()


@pytest.mark.parametrize("kw,msg", [
    (dict(prompt="a", height=484, width=832), "divisible by 8"),
    (dict(prompt="a", height=480, width=832, cb=["nope"]), "callback_on_step_end_tensor_inputs"),
    (dict(prompt="a", height=480, width=832, prompt_embeds=torch.zeros(1, 4, 8)), "Cannot forward both `prompt`"),
    (dict(prompt=None, height=480, width=832), "Provide either `prompt` or `prompt_embeds`"),
    (dict(prompt=3, height=480, width=832), "has to be of type `str` or `list`"),
    (dict(prompt="a", height=480, width=832, negative_prompt_embeds=torch.zeros(1, 4, 8)), "negative_prompt_embeds"),
    (dict(prompt=None, height=480, width=832, prompt_embeds=torch.zeros(1, 4, 8), negative_prompt="x",
          negative_prompt_embeds=torch.zeros(1, 4, 8)), "Cannot forward both `negative_prompt`"),
    (dict(prompt=None, height=480, width=832, prompt_embeds=torch.zeros(1, 4, 8), negative_prompt_embeds=torch.zeros(1, 5, 8)),
     "must have the same shape"),
    (dict(prompt="a", height=488, width=832), "divisible by 16"),
])
def test_check_inputs_raises_like_the_reference(kw, msg):
    """pipe.py:458-507: same conditions, same ValueError texts."""
    kw = dict(kw)
    pipe = _bare_pipeline()
    with pytest.raises(ValueError, match=msg.replace("`", ".")):
        pipe.check_inputs(kw.pop("prompt"), kw.pop("height"), kw.pop("width"), kw.pop("negative_prompt", None),
                          kw.pop("cb", ["latents"]), **kw)


def test_call_validates_before_touching_the_device():
    pipe = _bare_pipeline()
    with pytest.raises(ValueError, match="divisible by 8"):
        pipe(prompt="a", height=481, width=832)
    with pytest.raises(NotImplementedError, match="offload"):
        pipe.enable_model_cpu_offload(device="cuda")


def test_pipeline_to_moves_modules_and_returns_self():
    """`pipeline.to(device=device)` (inference.py:524)."""
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(2))
    from stableavatar_b200.pipeline import WanI2VTalkingInferenceLongPipeline
    mods = dict(text_encoder=M(), vae=M(), transformer=M(), clip_image_encoder=M())
    pipe = WanI2VTalkingInferenceLongPipeline(tokenizer=object(), wav2vec=M(), **mods)
    assert pipe.to(device="cpu") is pipe and pipe.to("cpu") is pipe
    pipe.to(torch.float64)
    assert pipe.transformer.p.dtype == torch.float64 and pipe.vae.p.dtype == torch.float32


def test_dsigma_falls_back_to_a_sigma_table():
    """A scheduler without the repo's dsigma_at (e.g. diffusers' own class, as inference.py passes) still works."""
    from stableavatar_b200.pipeline import _dsigma_at
    s = FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=5.0)
    s.set_timesteps(10, device="cpu", mu=1)

    class Foreign:
        sigmas = s.sigmas
    for i in range(10):
        assert _dsigma_at(Foreign(), i) == s.dsigma_at(i)


def test_fp8_weight_mode_gives_the_reference_values():
    """wan/utils/fp8_optimization.py:30-45 rounds every parameter except `modulation` to float8_e4m3fn and upcasts it
    for the forward; the B200 path keeps bf16 storage with exactly those values."""
    from stableavatar_b200 import synth
    from stableavatar_b200.fp8_optimization import convert_model_weight_to_float8, convert_weight_dtype_wrapper
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    cfg = dict(synth.DIT_TINY, num_layers=1)
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    m = WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys})
    sd = {k: v.bfloat16() for k, v in synth.dit_state_dict(cfg).items()}
    m.load_state_dict(sd, strict=True)
    m = m.to(torch.bfloat16)
    m._prep = "stale"
    convert_model_weight_to_float8(m, exclude_module_name=["modulation", ])          # inference.py:518
    convert_weight_dtype_wrapper(m, torch.bfloat16)
    assert m._prep is None
    for name, p in m.named_parameters():
        want = sd[name] if "modulation" in name else sd[name].to(torch.float8_e4m3fn).to(torch.bfloat16)
        assert p.dtype == torch.bfloat16 and torch.equal(p.data, want), name
    w = m.blocks[0].ffn[0].weight
    assert not torch.equal(w.data, sd["blocks.0.ffn.0.weight"])                      # the rounding is real
