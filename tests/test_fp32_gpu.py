"""GPU parity (-m gpu) of the fp32 mode (BASELINE config 1: fp32 weights; north-star tolerance 1e-4 per block): the
CUDA path with float32 parameters — every Linear / attention product a split-bf16 tensor-core GEMM — against the golden
fixtures written by the REAL reference in fp32 on CPU and against the fp32 oracle."""
import numpy as np
import pytest
import torch

from stableavatar_b200 import synth

pytestmark = pytest.mark.gpu
CFG = synth.DIT_TINY
TOL = 1e-4
SUB = (slice(None), slice(None), slice(0, None, 8))
KEYS = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
        "num_heads", "num_layers")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "dit_tiny.npz")


@pytest.fixture(scope="module")
def model():
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    m = WanTransformer3DFantasyModel(**{k: CFG[k] for k in KEYS})
    m.load_state_dict(synth.dit_state_dict(CFG), strict=True)
    return m.to("cuda", torch.float32)


def run(model, inp, **kw):
    model.hooks = {}
    dev = "cuda"
    out = model(x=inp["x"].to(dev), t=inp["t"].to(dev), context=[c.to(dev) for c in inp["context"]], seq_len=inp["seq_len"],
                clip_fea=inp["clip_fea"].to(dev), y=inp["y"].to(dev), vocal_embeddings=inp["vocal_embeddings"].to(dev),
                video_sample_n_frames=inp["video_sample_n_frames"], **kw)
    torch.cuda.synchronize()
    hooks, model.hooks = model.hooks, None
    return out, hooks


def test_split_gemm_is_fp32_accurate():
    """The building block: x w^T through the six-product split against a float64 product."""
    from stableavatar_b200 import fp32_mode as F
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(300, 1000, device="cuda", generator=g) * 3
    w = torch.randn(136, 1000, device="cuda", generator=g)
    b = torch.randn(136, device="cuda", generator=g)
    got = F.linear(x, w, b)
    want = x.double() @ w.double().t() + b.double()
    err, err_bf16 = rel(got, want), rel(x.bfloat16().float() @ w.bfloat16().float().t() + b, want)
    assert got.dtype == torch.float32 and err < 5e-6 and err < err_bf16 / 200, (err, err_bf16)
    tf = torch.nn.functional
    err = rel(F.linear(x, w, b, act=1), tf.gelu(want, approximate="tanh"))
    assert err < 5e-6, err


def test_fp32_attention_vs_float64():
    from stableavatar_b200 import fp32_mode as F
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v = (torch.randn(2, n, 3, 128, device="cuda", generator=g) for n in (70, 257, 257))
    got = F.attention(q, k, v)
    f = torch.nn.functional.scaled_dot_product_attention
    want = f(q.double().transpose(1, 2), k.double().transpose(1, 2), v.double().transpose(1, 2)).transpose(1, 2)
    assert rel(got, want) < 5e-6
    acc = F.attention(q, k, v, out=got.clone(), accumulate=True)
    assert rel(acc, 2 * want) < 5e-6


def test_cfg_batch_blocks_vs_reference_golden_1e4(model, gold):
    inp = synth.dit_inputs(CFG, frames=9, height=64, width=96)
    out, hooks = run(model, inp)
    assert out.dtype == torch.float32
    assert rel(hooks["vocal_context"][-1:], gold["A_vocal_context"]) < TOL       # the adapter ran once, on the last sample
    assert hooks["vocal_context"][0].abs().max().item() == 0
    for i in range(CFG["num_layers"]):
        assert rel(hooks[f"block{i}"][SUB], gold[f"A_block{i}"]) < TOL, i
    assert rel(out, gold["A_out"]) < TOL


def test_short_window_batch1_and_clip_level_vs_golden(model, gold):
    inp = synth.dit_inputs(CFG, frames=9, height=64, width=96, seed=1)             # live zero-pad tokens
    inp["x"], inp["y"] = inp["x"][:, :, :2].contiguous(), inp["y"][:, :, :2].contiguous()
    out, hooks = run(model, inp)
    assert rel(hooks["block1"][SUB], gold["B_block1"]) < TOL and rel(out, gold["B_out"]) < TOL
    out, _ = run(model, synth.dit_inputs(CFG, frames=5, height=64, width=64, batch=1, seed=2))
    assert rel(out, gold["C_out"]) < TOL
    out, _ = run(model, synth.dit_inputs(CFG, frames=9, height=64, width=96, seed=3), is_clip_level_modeling=True)
    assert rel(out, gold["D_out"]) < TOL


def test_14b_class_fp32_vs_golden(golden_dir):
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasy14BModel
    gold14 = np.load(golden_dir / "dit14b_tiny.npz")
    cfg = synth.DIT_14B_TINY
    m = WanTransformer3DFantasy14BModel(**{k: cfg[k] for k in KEYS})
    m.load_state_dict(synth.dit_state_dict(cfg), strict=True)
    m = m.to("cuda", torch.float32)
    inp = synth.dit_inputs(cfg, frames=81, height=32, width=32, batch=1, seed=2)
    m.hooks = {}
    out = m(x=inp["x"].cuda(), t=inp["t"].cuda(), context=[c.cuda() for c in inp["context"]], seq_len=inp["seq_len"],
            clip_fea=inp["clip_fea"].cuda(), y=inp["y"].cuda(), vocal_embeddings=inp["vocal_embeddings"].cuda())
    assert rel(m.hooks["block1"][:, :, ::4], gold14["C_block1"]) < TOL
    assert rel(out, gold14["C_out"]) < TOL


def test_fp32_denoise_step_vs_oracle():
    """One pipeline step (forward on the CFG batch + CFG + Euler) in fp32 against the oracle restatement."""
    from oracle import dit as O
    from stableavatar_b200.pipeline import WanI2VTalkingInferenceLongPipeline
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    sd = synth.dit_state_dict(CFG)
    m = WanTransformer3DFantasyModel(**{k: CFG[k] for k in KEYS})
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda", torch.float32)
    inp = synth.dit_inputs(CFG, frames=9, height=64, width=96, seed=7)
    pipe = WanI2VTalkingInferenceLongPipeline(transformer=m)
    lat = inp["x"][:1].cuda()
    new = pipe.denoise_step(lat, 900.0, -0.02, [c.cuda() for c in inp["context"]], inp["clip_fea"].cuda(), inp["y"].cuda(),
                            inp["vocal_embeddings"].cuda(), seq_len=inp["seq_len"], clip_length=9, text_guide_scale=3.0,
                            audio_guide_scale=5.0)
    with torch.no_grad():
        pred = O.dit_forward(sd, CFG, inp["x"], inp["t"], inp["context"], inp["seq_len"], inp["clip_fea"], inp["y"],
                             inp["vocal_embeddings"], 9)
    u, d, c = pred[0], pred[1], pred[2]
    want = inp["x"][0] + (-0.02) * (u + 5.0 * (d - u) + 3.0 * (c - d))
    assert new.dtype == torch.float32 and rel(new[0], want) < TOL


def test_config1_grid_480x832x5_fp32_full_width_vs_oracle():
    """BASELINE config 1 — 480x832, 5 frames (L = 2 x 30 x 52 = 3120), fp32, CFG batch 3 — at the full 1.3B width (dim 1536,
    12 heads, ffn 8960: 35 K-chunks in the FFN down projection) with 2 of the 30 layers, against the fp32 CPU oracle."""
    from oracle import dit as O
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    cfg = dict(synth.DIT_1_3B, num_layers=2, text_dim=256, text_len=32)
    sd = synth.dit_state_dict(cfg)
    m = WanTransformer3DFantasyModel(**{k: cfg[k] for k in KEYS})
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda", torch.float32)
    inp = synth.dit_inputs(cfg, frames=5, height=480, width=832, seed=7)
    out, hooks = run(m, inp)
    rh = {}
    with torch.no_grad():
        ref = O.dit_forward(sd, cfg, inp["x"], inp["t"], inp["context"], inp["seq_len"], inp["clip_fea"], inp["y"],
                            inp["vocal_embeddings"], 5, hooks=rh)
    assert out.shape == (3, 16, 2, 60, 104)
    for i in range(2):
        assert rel(hooks[f"block{i}"], rh[f"block{i}"]) < TOL, i
    assert rel(out, ref) < TOL
