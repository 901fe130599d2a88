"""GPU test (-m gpu, needs >= 2 GPUs, else skipped): the sequence-parallel forward equals the single-GPU forward (the
reference's single-GPU semantics are the oracle for SP, SURVEY.md fact #9-iii and §8c) — with the all-to-alls as direct
NVLink peer stores (csrc/sp_exchange.cu, the default), as peer stores pipelined per CFG sample, with the norm un-fused, and
over NCCL all_to_all_single (model.sp_exchange = "nccl"). Also the train_14B head count
(40 heads: pure Ulysses at P = 2 / 4 / 8, 20 heads per rank at P = 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stableavatar_b200 import synth

pytestmark = pytest.mark.gpu
CFG = synth.DIT_TINY


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


MODES = {"peer": dict(sp_exchange="peer"),                                      # the default: fused norm, serial order
         "peer_pipelined": dict(sp_exchange="peer", sp_fused_norm=True, sp_pipelined=True),
         "peer_serial_unfused": dict(sp_exchange="peer", sp_fused_norm=False, sp_pipelined=False, sp_fused_o=False),
         "nccl": dict(sp_exchange="nccl")}


def _build(arch):
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasy14BModel, WanTransformer3DFantasyModel
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    if arch == "1.3b":
        m = WanTransformer3DFantasyModel(**{k: CFG[k] for k in keys})
        m.load_state_dict({k: v.bfloat16() for k, v in synth.dit_state_dict(CFG).items()}, strict=True)
        m = m.to("cuda", torch.bfloat16)
        inp = synth.dit_inputs(CFG, frames=17, height=128, width=192, seed=5)      # L = 5*8*12 = 480
        return m, inp, dict(video_sample_n_frames=17)
    # the 14B head count at a CPU-free size: dim 5120 = 40 heads x 128, one layer, narrow FFN; weights drawn on the device
    cfg = dict(synth.DIT_14B, ffn_dim=1024, text_dim=128, text_len=24, num_layers=1)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            m = WanTransformer3DFantasy14BModel(**{k: cfg[k] for k in keys})
    finally:
        torch.set_default_dtype(old)
    m.init_random_(seed=3)
    inp = synth.dit_inputs(cfg, frames=81, height=64, width=64, seed=6)            # L = 21*4*4 = 336 = 8 * 42
    return m, inp, {}


def _worker(rank, world, port, q_out, mode, arch):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from stableavatar_b200 import ops
        ops.sp_set_barrier_timeout_ms(30_000)        # a protocol bug must not hang the box
        m, inp, extra = _build(arch)
        dev, bf = "cuda", torch.bfloat16
        kw = dict(x=inp["x"].to(dev, bf), t=inp["t"].to(dev), context=[c.to(dev, bf) for c in inp["context"]],
                  seq_len=inp["seq_len"], clip_fea=inp["clip_fea"].to(dev, bf), y=inp["y"].to(dev, bf),
                  vocal_embeddings=inp["vocal_embeddings"].to(dev, bf), **extra)
        single = m(**kw).float()
        m.enable_multi_gpus_inference()
        for k, v in MODES[mode].items():
            setattr(m, k, v)
        sp_out = m(**kw).float()
        again = m(**kw).float()                      # second call: receive buffers and barrier epochs are reused
        torch.cuda.synchronize()
        assert (getattr(m, "_sp_px", None) is not None) == (mode != "nccl")
        assert torch.equal(sp_out, again)
        q_out.put((rank, ((sp_out - single).norm() / single.norm()).item()))
    finally:
        dist.destroy_process_group()


def _run(world, mode, arch):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode, arch)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        for p in procs:
            p.join(240)
            assert p.exitcode == 0, f"rank process exit code {p.exitcode}"
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    res = dict(q.get(timeout=5) for _ in range(world))
    assert max(res.values()) < 1e-2, res
    return res


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sp_forward_equals_single_gpu(world, mode):
    """12 heads: pure Ulysses at P = 2 / 4, 4 head groups x 2 query splits at P = 8."""
    _run(world, mode, "1.3b")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sp_forward_40_heads_equals_single_gpu(world):
    """train_14B head count (wan/configs/wan_i2v_14B.py:26-35): 40 heads -> 20 / 10 / 5 heads per rank, pure Ulysses."""
    _run(world, "peer", "14b")
