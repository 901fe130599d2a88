"""GPU parity (-m gpu) of the Wan VAE decode and encode: CUDA path (bf16 tensor-core implicit-GEMM convs, channels-last) against
the fp32 CPU oracle and the golden fixtures of the real reference. Bars: rel-L2 <= 2e-2 (bf16 mode) and PSNR >= 35 dB
on the decoded frames (BASELINE.json), frames in [-1, 1] so peak-to-peak = 2."""
import math

import numpy as np
import pytest
import torch

from stableavatar_b200 import synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def psnr(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return 10 * math.log10(4.0 / ((a - b) ** 2).mean().item())


@pytest.fixture(scope="module")
def vae():
    from stableavatar_b200.wan_vae import AutoencoderKLWan
    m = AutoencoderKLWan()
    m.load_state_dict(synth.vae_state_dict(), strict=True)
    return m.to("cuda")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "vae_tiny.npz")


def test_three_latent_frames_vs_golden_and_oracle(vae, gold):
    from oracle import vae as V
    z = synth.det_normal("vae_z", (1, 16, 3, 6, 8))
    out = vae.decode(z.cuda()).sample
    torch.cuda.synchronize()
    assert out.shape == (1, 3, 9, 48, 64) and out.dtype == torch.float32
    assert out.abs().max().item() <= 1.0
    with torch.no_grad():
        ref = V.vae_decode(synth.vae_state_dict(), z)
    assert rel(out, ref) < 2e-2 and psnr(out, ref) > 35
    assert rel(out, gold["z3_out"]) < 2e-2 and psnr(out, gold["z3_out"]) > 35


def test_single_frame_batch(vae, gold):
    z = synth.det_normal("vae_z1", (2, 16, 1, 4, 6))
    out = vae.decode(z.cuda(), return_dict=False)[0]
    assert out.shape == (2, 3, 1, 32, 48)
    assert rel(out, gold["z1_out"]) < 2e-2 and psnr(out, gold["z1_out"]) > 35


def test_ragged_spatial_tiles_vs_oracle(vae):
    """h, w not multiples of the 16 x 8 output tile at any stage: masked stores and TMA zero-fill halos."""
    from oracle import vae as V
    z = synth.det_normal("vae_z2", (1, 16, 2, 5, 7))
    out = vae.decode(z.cuda()).sample
    with torch.no_grad():
        ref = V.vae_decode(synth.vae_state_dict(), z)
    assert out.shape == ref.shape == (1, 3, 5, 40, 56)
    assert rel(out, ref) < 2e-2 and psnr(out, ref) > 35


def test_opt_in_graph_replay_of_steady_state_chunks_is_bit_identical(vae):
    """use_cuda_graph: chunk 2 of the second decode of a shape is captured and replayed for chunks 2.. (and for every later
    decode of that shape); the frames must not change, also not for another latent decoded through the cached graph."""
    z1 = synth.det_normal("vae_zg1", (1, 16, 6, 12, 16)).cuda()
    z2 = synth.det_normal("vae_zg2", (1, 16, 6, 12, 16)).cuda()
    want1, want2 = vae.decode(z1).sample.clone(), vae.decode(z2).sample.clone()
    vae.use_cuda_graph = True
    try:
        vae.decode(z1)                                   # shape seen once: eager
        got1 = vae.decode(z1).sample.clone()             # captures at chunk 2, replays the rest
        assert vae._dec_graph is not None
        got2 = vae.decode(z2).sample.clone()             # replays the cached graph on another latent
    finally:
        vae.use_cuda_graph = False
        vae._dec_graph = vae._dec_seen = None
    assert torch.equal(got1, want1) and torch.equal(got2, want2)


def test_decode_is_idempotent_across_calls(vae):
    """The ring-buffer caches are reset per decode: the same latent decodes to the same frames twice."""
    z = synth.det_normal("vae_z", (1, 16, 3, 6, 8)).cuda()
    a = vae.decode(z).sample.clone()
    b = vae.decode(z).sample
    assert torch.equal(a, b)


def test_benchmark_grid_vs_real_reference_subsample(vae, golden_dir):
    """SURVEY.md §8c pin #4: the decode at the BENCHMARK's latent grid — z [1,16,3,60,104] -> [1,3,9,480,832], the conv
    shapes of config 4 (96->96 @480x832, 192->192 @240x416 ...) — against the REAL reference's output, kept as every 8th
    row / column (tools/gen_golden_vae.py gen_vae_fullres) plus per-frame first and second moments of all pixels."""
    gold = np.load(golden_dir / "vae_fullres_sub.npz")
    z = synth.det_normal("vae_z_full", (1, 16, 3, 60, 104))
    out = vae.decode(z.cuda()).sample
    torch.cuda.synchronize()
    assert tuple(out.shape) == (1, 3, 9, 480, 832)
    sub, ref = out[..., 3::8, 5::8], gold["sub8"].astype(np.float32)
    assert rel(sub, ref) < 2e-2 and psnr(sub, ref) > 35, (rel(sub, ref), psnr(sub, ref))
    assert np.allclose(out.mean(dim=(-1, -2)).cpu().numpy(), gold["frame_mean"], atol=5e-3)   # bf16 activations: per-frame mean within 0.25 % of the [-1, 1] range
    assert np.allclose((out.double() ** 2).mean(dim=(-1, -2)).cpu().numpy(), gold["frame_sq"], rtol=3e-2, atol=1e-4)


def _pp_worker(rank, world, port, q_out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from stableavatar_b200.wan_vae import AutoencoderKLWan
        m = AutoencoderKLWan()
        m.load_state_dict(synth.vae_state_dict(), strict=True)
        m = m.to(torch.device("cuda", rank))
        z = synth.det_normal("vae_z", (1, 16, 3, 6, 8)).cuda()
        single = m.decode(z).sample.clone()
        m.enable_multi_gpus_decode()
        pp = m.decode(z).sample
        torch.cuda.synchronize()
        q_out.put((rank, bool(torch.equal(pp, single))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_pipeline_parallel_decode_equals_single_gpu(world):
    """Depth-pipelined multi-GPU decode: same ops, same caches -> bit-identical frames on every rank."""
    import socket
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()), res


# ---------------------------------------------------------------------------------------------- encode
@pytest.fixture(scope="module")
def vae_enc():
    from stableavatar_b200.wan_vae import AutoencoderKLWan
    m = AutoencoderKLWan()
    m.load_state_dict(synth.vae_state_dict(encoder=True), strict=True)
    return m.to("cuda")


@pytest.fixture(scope="module")
def enc_gold(golden_dir):
    return np.load(golden_dir / "vae_enc_tiny.npz")


@pytest.mark.parametrize("name,shape", [("x9", (1, 3, 9, 32, 48)), ("x1", (2, 3, 1, 16, 32)), ("x6", (1, 3, 6, 16, 16))])
def test_encode_vs_golden(vae_enc, enc_gold, name, shape):
    """Chunks 1, 4, 4 (both stride-2 time convs with their one-frame caches), the first-chunk-only path with batch 2,
    and a clip whose trailing frame the 1 + 4k chunking drops — against the real reference's outputs."""
    x = synth.det_normal("vae_" + name, shape).clamp_(-1, 1)
    post = vae_enc.encode(x.cuda())[0]                      # the reference call pattern: vae.encode(x)[0].mode()
    mode = post.mode()
    torch.cuda.synchronize()
    ref = enc_gold[name + "_mode"]
    assert tuple(mode.shape) == ref.shape and mode.dtype == torch.float32
    assert rel(mode, ref) < 2e-2
    if name == "x9":
        assert rel(post.parameters, enc_gold["x9_params"]) < 2e-2
        assert torch.equal(vae_enc.encode(x.cuda()).latent_dist.mode(), mode)      # caches reset between calls


def test_encode_ragged_tiles_vs_oracle(vae_enc):
    """H/8, W/8 not multiples of the conv tile at any stage; 13 frames = chunks 1, 4, 4, 4."""
    from oracle import vae as V
    x = synth.det_normal("vae_x13", (1, 3, 13, 40, 56)).clamp_(-1, 1)
    out = vae_enc.encode(x.cuda(), return_dict=False)[0].parameters
    with torch.no_grad():
        ref = V.vae_encode(synth.vae_state_dict(encoder=True), x)
    assert out.shape == ref.shape == (1, 32, 4, 5, 7)
    assert rel(out, ref) < 2e-2


def test_encode_then_decode_shapes_and_decode_only_state_dict(vae_enc, vae):
    x = synth.det_normal("vae_x5", (1, 3, 5, 32, 32)).clamp_(-1, 1)
    z = vae_enc.encode(x.cuda()).latent_dist.mode()
    assert z.shape == (1, 16, 2, 4, 4)
    assert vae_enc.decode(z).sample.shape == (1, 3, 5, 32, 32)
    with pytest.raises(RuntimeError, match="encoder weights"):
        vae.encode(x.cuda())                                 # `vae` was loaded from a decode-only state dict
    with pytest.raises(ValueError, match="multiples of 8"):
        vae_enc.encode(torch.zeros(1, 3, 1, 20, 32, device="cuda"))
