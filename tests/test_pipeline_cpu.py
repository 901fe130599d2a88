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
