"""Generate tests/golden/*.npz by running the REAL reference modules (imported from /root/reference through
oracle/refstub.py) on the deterministic synthetic weights/inputs of stableavatar_b200/synth.py.

Run in the build container only (the GPU box has no /root/reference):   python tools/gen_golden.py [dit|vae|all]
The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.refstub import import_reference  # noqa: E402
from stableavatar_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
SUB = (slice(None), slice(None), slice(0, None, 8))          # channel subsample for per-block tensors


def build_ref_dit(dit, cfg):
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    m = dit.WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys}).eval()
    m.load_state_dict(synth.dit_state_dict(cfg), strict=True)
    return m


def run_ref(m, inp, **kw):
    blocks = {}
    hs = [b.register_forward_hook(lambda mod, a, out, i=i: blocks.__setitem__(f"block{i}", out.detach().clone()))
          for i, b in enumerate(m.blocks)]
    hs.append(m.vocal_projector.register_forward_hook(
        lambda mod, a, out: blocks.__setitem__("vocal_context", out[0].detach().clone())))
    with torch.no_grad():
        out = m(x=inp["x"], t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], clip_fea=inp["clip_fea"],
                y=inp["y"], vocal_embeddings=inp["vocal_embeddings"],
                video_sample_n_frames=inp["video_sample_n_frames"], **kw)
    for h in hs:
        h.remove()
    return out, blocks


def gen_dit():
    dit, vp, _ = import_reference()
    from wan.models.vocal_projector_fantasy import split_audio_sequence, split_tensor_with_padding
    cfg = synth.DIT_TINY
    m = build_ref_dit(dit, cfg)
    res = {}

    # case A: CFG batch of 3 (pipe.py:730-750), 9 frames @ 64x96 -> 3 latent frames of 4x6 tokens
    inp = synth.dit_inputs(cfg, frames=9, height=64, width=96)
    out, blocks = run_ref(m, inp)
    res["A_out"] = out.numpy()
    for k, v in blocks.items():
        res["A_" + k] = v.numpy() if k == "vocal_context" else v[SUB].numpy()
        res["A_" + k + "_norm"] = np.array(v.double().norm().item())

    # case B: short last sliding window: 2 latent frames but seq_len / audio split for 3
    inp = synth.dit_inputs(cfg, frames=9, height=64, width=96, seed=1)
    inp["x"], inp["y"] = inp["x"][:, :, :2].contiguous(), inp["y"][:, :, :2].contiguous()
    out, blocks = run_ref(m, inp)
    res["B_out"] = out.numpy()
    res["B_block1"] = blocks["block1"][SUB].numpy()

    # case C: batch 1 (no CFG): the adapter runs on the sample itself (1B.py:1008-1009)
    inp = synth.dit_inputs(cfg, frames=5, height=64, width=64, batch=1, seed=2)
    out, blocks = run_ref(m, inp)
    res["C_out"] = out.numpy()

    # case D: clip-level audio modelling (1B.py:1011-1015, 587-596)
    inp = synth.dit_inputs(cfg, frames=9, height=64, width=96, seed=3)
    out, _ = run_ref(m, inp, is_clip_level_modeling=True)
    res["D_out"] = out.numpy()

    # case E: TeaCache decisions + outputs over 6 steps (1B.py:1021-1103), threshold chosen so some steps skip
    inp = synth.dit_inputs(cfg, frames=9, height=64, width=96, seed=4)
    coeff = [-5.21862437e+04, 9.23041404e+03, -5.28275948e+02, 1.36987616e+01, -4.99875664e-02]
    m.enable_teacache(coeff, num_steps=6, rel_l1_thresh=0.15, num_skip_start_steps=1, offload=False)
    ts = [999.0, 960.0, 920.0, 870.0, 800.0, 700.0]
    outs, calc = [], []
    for tv in ts:
        inp["t"] = torch.full((3,), tv)
        o, _ = run_ref(m, inp)
        outs.append(o.numpy())
        calc.append(bool(m.teacache.should_calc))
    m.disable_teacache()
    res["E_t"] = np.array(ts, dtype=np.float32)
    res["E_out"] = np.stack(outs)
    res["E_should_calc"] = np.array(calc)

    # audio window tables (vp.py:39-131)
    for T, nf in ((9, 5), (17, 9), (134, 81), (161, 81), (173, 81), (161, 69)):
        r = split_audio_sequence(T, num_frames=nf)
        sub, lens = split_tensor_with_padding(torch.arange(1, T + 1, dtype=torch.float32).view(1, T, 1), r, 4)
        res[f"win_{T}_{nf}_ranges"] = np.array(r, dtype=np.int64)
        res[f"win_{T}_{nf}_gather"] = sub[0, :, :, 0].numpy().astype(np.int64)     # 1-based source index, 0 = pad
        res[f"win_{T}_{nf}_lens"] = lens.numpy()

    # RoPE table and sinusoid (fp64 islands)
    fr = m.freqs
    res["rope_freqs_real"] = fr.real[:64].numpy()
    res["rope_freqs_imag"] = fr.imag[:64].numpy()
    res["sinusoid"] = dit.sinusoidal_embedding_1d(256, torch.tensor([0.0, 1.0, 500.5, 999.0])).numpy()
    q = synth.det_normal("rope_q", (2, 50, 3, 128))
    res["rope_apply"] = dit.rope_apply(q, torch.tensor([[2, 4, 6], [2, 4, 6]]), fr).numpy()

    np.savez_compressed(GOLD / "dit_tiny.npz", **res)
    print("wrote", GOLD / "dit_tiny.npz", sum(v.nbytes for v in res.values()) / 1e6, "MB raw")


def gen_dit_14b():
    """train_14B architecture at a CPU-sized width (dim 512 = 4 heads of 128, adapter 8 heads of 64): the real
    WanTransformer3DFantasy14BModel on 81 frames @ 32x32 (21 latent frames of 2x2 tokens — the class hard-codes 21)."""
    from oracle.refstub import import_reference_14b
    dit14, _ = import_reference_14b()
    cfg = synth.DIT_14B_TINY
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    m = dit14.WanTransformer3DFantasy14BModel(**{k: cfg[k] for k in keys}).eval()
    m.load_state_dict(synth.dit_state_dict(cfg), strict=True)
    res = {}
    for tag, kw in (("A", dict(batch=3)), ("C", dict(batch=1, seed=2))):
        inp = synth.dit_inputs(cfg, frames=81, height=32, width=32, **kw)
        blocks = {}
        hs = [b.register_forward_hook(lambda mod, a, out, i=i: blocks.__setitem__(f"block{i}", out.detach().clone()))
              for i, b in enumerate(m.blocks)]
        hs.append(m.vocal_projector.register_forward_hook(
            lambda mod, a, out: blocks.__setitem__("vocal_context", out[0].detach().clone())))
        with torch.no_grad():
            out = m(x=inp["x"], t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], clip_fea=inp["clip_fea"],
                    y=inp["y"], vocal_embeddings=inp["vocal_embeddings"])
        for h in hs:
            h.remove()
        res[tag + "_out"] = out.numpy()
        for k, v in blocks.items():
            res[f"{tag}_{k}"] = v[:, :, ::4].numpy()                       # every 4th channel
    np.savez_compressed(GOLD / "dit14b_tiny.npz", **res)
    print("wrote", GOLD / "dit14b_tiny.npz", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_grad_enabled(False)
    if what in ("dit", "all"):
        gen_dit()
    if what in ("dit14b", "all"):
        gen_dit_14b()
    if what in ("vae", "all"):
        from tools.gen_golden_vae import gen_vae
        gen_vae()
