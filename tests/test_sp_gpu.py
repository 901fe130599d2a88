"""GPU test (-m gpu, needs >= 2 GPUs, else skipped): the sequence-parallel forward equals the single-GPU forward (the
reference's single-GPU semantics are the oracle for SP, SURVEY.md fact #9-iii and §8c) — with the all-to-alls as direct
NVLink peer stores (csrc/sp_exchange.cu, the default) and over NCCL all_to_all_single (SA_SP_PEER=0)."""
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


def _worker(rank, world, port, q_out, mode):
    peer, fused = mode[0], mode[1]          # "11": peer stores with fused norm+rope, "10": peer stores, "00": NCCL
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SA_SP_PEER=peer, SA_SP_FUSED_NORM=fused)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
        keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
                "num_heads", "num_layers")
        m = WanTransformer3DFantasyModel(**{k: CFG[k] for k in keys})
        m.load_state_dict({k: v.bfloat16() for k, v in synth.dit_state_dict(CFG).items()}, strict=True)
        m = m.to("cuda", torch.bfloat16)
        inp = synth.dit_inputs(CFG, frames=17, height=128, width=192, seed=5)      # L = 5*8*12 = 480
        dev, bf = "cuda", torch.bfloat16
        kw = dict(x=inp["x"].to(dev, bf), t=inp["t"].to(dev), context=[c.to(dev, bf) for c in inp["context"]],
                  seq_len=inp["seq_len"], clip_fea=inp["clip_fea"].to(dev, bf), y=inp["y"].to(dev, bf),
                  vocal_embeddings=inp["vocal_embeddings"].to(dev, bf), video_sample_n_frames=17)
        single = m(**kw).float()
        m.enable_multi_gpus_inference()
        sp_out = m(**kw).float()
        again = m(**kw).float()                      # second call: receive buffers and barrier epochs are reused
        torch.cuda.synchronize()
        assert (getattr(m, "_sp_px", None) is not None) == (peer == "1")
        assert torch.equal(sp_out, again)
        q_out.put((rank, ((sp_out - single).norm() / single.norm()).item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("peer", ["11", "10", "00"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sp_forward_equals_single_gpu(world, peer):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, peer)) for r in range(world)]
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
