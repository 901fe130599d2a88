"""Pin the CPU oracle (oracle/dit.py) against fixtures produced by the REAL reference modules
(tests/golden/dit_tiny.npz, written by tools/gen_golden.py from /root/reference). fp32 on CPU: tolerance 1e-4
relative L2 (BASELINE.json fp32-mode bar); index tables must match exactly."""
import numpy as np
import pytest
import torch

from oracle import dit as O
from stableavatar_b200 import synth

CFG = synth.DIT_TINY
SUB = (slice(None), slice(None), slice(0, None, 8))


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "dit_tiny.npz")


@pytest.fixture(scope="module")
def sd():
    return synth.dit_state_dict(CFG)


def run(sd, inp, **kw):
    hooks = {}
    with torch.no_grad():
        out = O.dit_forward(sd, CFG, inp["x"], inp["t"], inp["context"], inp["seq_len"], inp["clip_fea"], inp["y"],
                            inp["vocal_embeddings"], inp["video_sample_n_frames"], hooks=hooks, **kw)
    return out, hooks


def test_cfg_batch_forward_and_blocks(gold, sd):
    inp = synth.dit_inputs(CFG, frames=9, height=64, width=96)
    out, hooks = run(sd, inp)
    assert rel(out, gold["A_out"]) < 1e-4
    assert rel(hooks["vocal_context"][1:2], gold["A_vocal_context"]) < 1e-4
    assert hooks["vocal_context"][0].abs().max() == 0          # uncond sample sees an all-zero audio context
    for i in range(CFG["num_layers"]):
        assert rel(hooks[f"block{i}"][SUB], gold[f"A_block{i}"]) < 1e-4
        assert abs(hooks[f"block{i}"].double().norm().item() / gold[f"A_block{i}_norm"] - 1) < 1e-5


def test_short_window_live_pad_tokens(gold, sd):
    inp = synth.dit_inputs(CFG, frames=9, height=64, width=96, seed=1)
    inp["x"], inp["y"] = inp["x"][:, :, :2].contiguous(), inp["y"][:, :, :2].contiguous()
    out, hooks = run(sd, inp)
    assert out.shape == (3, 16, 2, 8, 12)
    assert rel(out, gold["B_out"]) < 1e-4
    assert rel(hooks["block1"][SUB], gold["B_block1"]) < 1e-4


def test_batch1_and_clip_level(gold, sd):
    inp = synth.dit_inputs(CFG, frames=5, height=64, width=64, batch=1, seed=2)
    assert rel(run(sd, inp)[0], gold["C_out"]) < 1e-4
    inp = synth.dit_inputs(CFG, frames=9, height=64, width=96, seed=3)
    assert rel(run(sd, inp, is_clip_level_modeling=True)[0], gold["D_out"]) < 1e-4


def test_seq_len_not_divisible_by_groups_raises(sd):
    inp = synth.dit_inputs(CFG, frames=9, height=64, width=96)
    inp["x"], inp["y"] = inp["x"][:, :, :2].contiguous(), inp["y"][:, :, :2].contiguous()
    inp["seq_len"] = 2 * 4 * 6 + 1                               # 49 tokens cannot be viewed as 3 groups
    with pytest.raises(RuntimeError):
        run(sd, inp)


def test_teacache_sequence(gold, sd):
    inp = synth.dit_inputs(CFG, frames=9, height=64, width=96, seed=4)
    coeff = [-5.21862437e+04, 9.23041404e+03, -5.28275948e+02, 1.36987616e+01, -4.99875664e-02]
    tc = O.TeaCache(coeff, num_steps=6, rel_l1_thresh=0.15, num_skip_start_steps=1)
    assert not gold["E_should_calc"].all() and gold["E_should_calc"].any()
    for i, tv in enumerate(gold["E_t"]):
        inp["t"] = torch.full((3,), float(tv))
        cnt_before = tc.cnt
        out, hooks = run(sd, inp, teacache=tc)
        assert ("block0" in hooks) == bool(gold["E_should_calc"][i]), (i, cnt_before)
        assert rel(out, gold["E_out"][i]) < 1e-4


@pytest.mark.parametrize("T,nf", [(9, 5), (17, 9), (134, 81), (161, 81), (173, 81), (161, 69)])
def test_audio_windows(gold, T, nf):
    r = O.split_audio_sequence(T, num_frames=nf)
    assert np.array_equal(np.array(r), gold[f"win_{T}_{nf}_ranges"])
    sub, lens = O.split_tensor_with_padding(torch.arange(1, T + 1, dtype=torch.float32).view(1, T, 1), r, 4)
    assert np.array_equal(sub[0, :, :, 0].numpy().astype(np.int64), gold[f"win_{T}_{nf}_gather"])
    assert np.array_equal(lens.numpy(), gold[f"win_{T}_{nf}_lens"])


def test_audio_window_docstring_example():
    # vp.py:88 docstring: split_audio_sequence(173) starts [[-7, 1], [1, 9], ...]
    r = O.split_audio_sequence(173, num_frames=81)
    assert r[0] == [-7, 1] and r[1] == [1, 9] and len(r) == 21


def test_rope_and_sinusoid(gold):
    fr = O.rope_freqs(128)
    assert np.allclose(fr.real[:64].numpy(), gold["rope_freqs_real"], atol=1e-12)
    assert np.allclose(fr.imag[:64].numpy(), gold["rope_freqs_imag"], atol=1e-12)
    s = O.sinusoidal_embedding_1d(256, torch.tensor([0.0, 1.0, 500.5, 999.0]))
    assert np.allclose(s.numpy(), gold["sinusoid"], atol=1e-12)
    q = synth.det_normal("rope_q", (2, 50, 3, 128))
    out = O.rope_apply(q, [(2, 4, 6), (2, 4, 6)], fr)
    assert np.allclose(out.numpy(), gold["rope_apply"], atol=1e-6)
    assert torch.equal(out[:, 48:], q[:, 48:])                   # tokens beyond f*h*w are not rotated


# ---------------------------------------------------------------------------------------------- Wan VAE decode
@pytest.fixture(scope="module")
def vae_gold(golden_dir):
    return np.load(golden_dir / "vae_tiny.npz")


def test_vae_decode_three_latent_frames(vae_gold):
    """1 + 4 + 4 frames: first-chunk 'Rep' path, 1-frame caches growing to 2, both temporal upsamplers."""
    from oracle import vae as V
    sd = synth.vae_state_dict()
    with torch.no_grad():
        out = V.vae_decode(sd, synth.det_normal("vae_z", (1, 16, 3, 6, 8)))
    assert out.shape == (1, 3, 9, 48, 64)
    assert rel(out, vae_gold["z3_out"]) < 1e-4
    assert out.abs().max() <= 1.0


def test_vae_decode_single_frame_batch(vae_gold):
    from oracle import vae as V
    sd = synth.vae_state_dict()
    with torch.no_grad():
        out = V.vae_decode(sd, synth.det_normal("vae_z1", (2, 16, 1, 4, 6)))
    assert out.shape == (2, 3, 1, 32, 48)
    assert rel(out, vae_gold["z1_out"]) < 1e-4


# ---------------------------------------------------------------------------------------------- Wan VAE encode
@pytest.fixture(scope="module")
def vae_enc_gold(golden_dir):
    return np.load(golden_dir / "vae_enc_tiny.npz")


@pytest.mark.parametrize("name,shape,lat", [("x9", (1, 3, 9, 32, 48), (1, 16, 3, 4, 6)),       # chunks 1, 4, 4
                                            ("x1", (2, 3, 1, 16, 32), (2, 16, 1, 2, 4)),       # first-chunk path only, batch 2
                                            ("x6", (1, 3, 6, 16, 16), (1, 16, 2, 2, 2))])      # trailing frame dropped
def test_vae_encode_vs_reference(vae_enc_gold, name, shape, lat):
    from oracle import vae as V
    sd = synth.vae_state_dict(encoder=True)
    with torch.no_grad():
        out = V.vae_encode(sd, synth.det_normal("vae_" + name, shape).clamp_(-1, 1))
    assert out.shape == (lat[0], 32, *lat[2:])
    assert rel(out[:, :16], vae_enc_gold[name + "_mode"]) < 1e-5
    if name == "x9":
        assert rel(out, vae_enc_gold["x9_params"]) < 1e-5


def test_vae_state_dict_decoder_values_do_not_depend_on_encoder_flag():
    a, b = synth.vae_state_dict(), synth.vae_state_dict(encoder=True)
    assert set(a) < set(b) and all(torch.equal(a[k], b[k]) for k in a)


# ---------------------------------------------------------------------------------------------- train_14B architecture
@pytest.mark.parametrize("tag,kw", [("A", dict(batch=3)), ("C", dict(batch=1, seed=2))])
def test_14b_architecture_vs_reference(golden_dir, tag, kw):
    """WanTransformer3DFantasy14BModel at a CPU-sized width: two-stage audio projection, adapter at the DiT width run on
    every sample of the batch (no [0, vc, vc] replication), 21 audio groups hard-coded."""
    gold = np.load(golden_dir / "dit14b_tiny.npz")
    cfg = synth.DIT_14B_TINY
    sd14 = synth.dit_state_dict(cfg)
    inp = synth.dit_inputs(cfg, frames=81, height=32, width=32, **kw)
    hooks = {}
    with torch.no_grad():
        out = O.dit_forward(sd14, cfg, inp["x"], inp["t"], inp["context"], inp["seq_len"], inp["clip_fea"], inp["y"],
                            inp["vocal_embeddings"], hooks=hooks)
    assert rel(out, gold[tag + "_out"]) < 1e-5
    assert rel(hooks["vocal_context"][:, :, ::4], gold[tag + "_vocal_context"]) < 1e-5
    for i in range(2):
        assert rel(hooks[f"block{i}"][:, :, ::4], gold[f"{tag}_block{i}"]) < 1e-5
    if tag == "A":                                  # sample 0 gets the adapter's answer to silent audio, not zeros
        assert hooks["vocal_context"][0].abs().max() > 0
