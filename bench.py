#!/usr/bin/env python
"""bench.py — seconds per denoise step of the StableAvatar 1.3B audio-DiT at 480x832x81 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference algorithm (oracle port) on host cores

A "step" is one denoise step of the pipeline loop body (wan/pipeline/wan_inference_long_pipeline.py:730-754): the DiT
forward on the CFG batch of 3 (text/audio drop-outs), the 3-way CFG combine and the flow-matching Euler update.
Prints ONE JSON line on rank 0. Weights are random-init of the named architecture and inputs synthetic (no network).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "s/denoise-step 1.3B @480x832x81f"
SELF_ATTN_DRAM_BYTES_B3 = 913.93e6 + 286.01e6   # ncu --set full of one flash_attn_v8_kernel launch inside bench.py (profiles/r02_selfattn_in_bench_ncu.txt)
UNIT = "s/step"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(bf16=d["bf16_tflops_sustained"], bf16_burst=d["bf16_tflops"], hbm=d["hbm_gbs"], src="measured")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def nvlink_counters(index):
    """Sum of the NVLink data counters of one GPU in bytes (nvidia-smi nvlink -gt d), or None where unsupported."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=10).stdout
    except Exception:
        return None
    import re
    tx = [int(m) for m in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
    rx = [int(m) for m in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
    if not tx or not rx:
        return None
    return sum(tx) * 1024, sum(rx) * 1024


def workload(args):
    F_lat, h, w = (args.frames - 1) // 4 + 1, args.height // 8, args.width // 8
    L = F_lat * (h // 2) * (w // 2)
    return F_lat, h, w, L


def step_flops(cfg, L, B=3, text=512, clip=257, G=21, A=15):
    """SURVEY.md §8d algorithmic FLOPs of one denoise step (no padding, no recompute)."""
    d, f, nl = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"]
    gemm = 2 * L * d * (6 * d + 2 * f) + 4 * d * d * (text + clip + G * A)
    self_attn = 4 * L * L * d
    cross = 4 * L * (text + clip + A) * d
    adapter = 2 * (2 * 2 * L * d * d) + 2 * 4 * G * A * L // G * d
    return B * nl * (gemm + self_attn + cross) + adapter, B * self_attn


# ---------------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    """All host threads for the CPU arm: torchrun exports OMP_NUM_THREADS=1, which would make the reference arm 12x slower
    than the same code launched directly (round-1 SCALE ratios were inflated by exactly that)."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except AttributeError:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_block_case(cfg, L, grid):
    """Weights and inputs of ONE WanAttentionBlock at the full sequence length, B=1 (bf16-representable values, so the
    B200 block can be run on exactly the same numbers)."""
    from oracle import dit as O
    from stableavatar_b200 import synth
    one = dict(cfg, num_layers=1)
    shapes = {k: v for k, v in synth.dit_param_shapes(one).items() if k.startswith("blocks.0.")}
    g = torch.Generator().manual_seed(0)
    r = lambda t: t.bfloat16().float()  # noqa: E731
    sd = {k: r(torch.randn(v, generator=g) * (0.02 if len(v) < 2 or k.endswith("bias") else v[-1] ** -0.5)) for k, v in shapes.items()}
    for k in list(sd):
        if k.endswith("norm_q.weight") or k.endswith("norm_k.weight") or k.endswith("norm_k_img.weight") or k.endswith("norm3.weight"):
            sd[k] = r(1.0 + 5.0 * sd[k])                      # norm scales around 1
    d = cfg["dim"]
    G = grid[0]
    return dict(sd=sd, x=r(torch.randn(1, L, d, generator=g)), e0=r(torch.randn(1, 6, d, generator=g) * 0.1),
                ctx=r(torch.randn(1, 257 + 512, d, generator=g)), vc=r(torch.randn(1, G, 15, d, generator=g)), grid=grid, G=G,
                freqs=O.rope_freqs(d // cfg["num_heads"]), heads=cfg["num_heads"])


def cpu_block_run(case):
    """Time the oracle block (CPU restatement of 1B.py:650-695) on `case`; returns (seconds, output [1, L, C])."""
    from oracle import dit as O
    t0 = time.perf_counter()
    with torch.no_grad():
        out = O.dit_block(case["sd"], "blocks.0.", case["x"], case["e0"], [case["grid"]], case["freqs"], case["ctx"], case["vc"],
                          case["G"], case["heads"])
    return time.perf_counter() - t0, out


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores. The Python reference cannot travel to the GPU box
    (no /root/reference there), so this is the oracle port (oracle/dit.py, pinned to the real reference by
    tests/golden). Each step = one block at full L, B=1, extrapolated x layers x CFG batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from stableavatar_b200 import synth
    cfg = synth.DIT_1_3B
    F_lat, h, w, L = workload(args)
    cores = host_threads()
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    case = cpu_block_case(cfg, L, (F_lat, h // 2, w // 2))
    for _ in range(warm):
        cpu_block_run(case)
    ts = [cpu_block_run(case)[0] for _ in range(steps)]
    per_block = sum(ts) / len(ts)
    value = per_block * cfg["num_layers"] * 3
    sample = (f"1 WanAttentionBlock (oracle port, fp32) at L={L}, B=1: {per_block:.2f} s measured on {cores} threads; "
              f"x{cfg['num_layers']} blocks x3 CFG samples extrapolated; {steps} timed / {warm} warm-up samples")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, L)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- B200 arm
def build_model(cfg, device):
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasy14BModel, WanTransformer3DFantasyModel
    if cfg.get("variant") == "14B":
        WanTransformer3DFantasyModel = WanTransformer3DFantasy14BModel      # noqa: N806  (train_14B architecture, config 5)
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            m = WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys})
    finally:
        torch.set_default_dtype(old)
    return m.init_random_(seed=0)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()


def gpu_block_parity(cfg, case, dev):
    """The B200 WanAttentionBlock at the benchmarked size (dim 1536, ffn 8960, L = 32 760, B = 1) on the weights and inputs
    of the CPU-baseline sample -> [1, L, C] for the comparison with the oracle output (north_star: <= 2e-2 per block)."""
    one = dict(cfg, num_layers=1)
    m = build_model(one, dev)
    own = m.state_dict()
    own.update({k: v.to(dev, torch.bfloat16) for k, v in case["sd"].items()})
    m.load_state_dict(own, strict=True)
    out = m.block_forward(0, case["x"].to(dev), case["e0"].to(dev), case["ctx"].to(dev), case["vc"].to(dev), case["grid"])
    torch.cuda.synchronize()
    return out.float().cpu()


def run_b200(args):
    import torch.distributed as dist
    from stableavatar_b200 import _lib as L_, ops, synth
    from stableavatar_b200.pipeline import WanI2VTalkingInferenceLongPipeline
    from stableavatar_b200.scheduler import FlowMatchEulerDiscreteScheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ops.sp_set_barrier_timeout_ms(120_000)      # the ranks of a bench run enter every forward together
    cfg = synth.DIT_14B if args.model == "14b" else synth.DIT_1_3B
    F_lat, h, w, L = workload(args)
    model = build_model(cfg, dev)
    sched = FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=5.0)
    sched.set_timesteps(50, device=dev)
    pipe = WanI2VTalkingInferenceLongPipeline(transformer=model, scheduler=sched)

    # synthetic conditioning of the pipeline's shapes, first in pinned host memory
    inp = synth.dit_inputs(cfg, frames=args.frames, height=args.height, width=args.width, text_tokens=64)
    bf = torch.bfloat16
    Wn = args.windows                                    # windows of one sliding-window step batched into the forward
    host = dict(latents=inp["x"][:1].to(bf).repeat(Wn, 1, 1, 1, 1), y=inp["y"].to(bf), clip=inp["clip_fea"].to(bf),
                audio=inp["vocal_embeddings"].to(bf).repeat(Wn, 1, 1), ctx=[c.to(bf) for c in inp["context"]])
    host = {k: ([t.pin_memory() for t in v] if isinstance(v, list) else v.pin_memory()) for k, v in host.items()}
    h2d_bytes = sum(t.numel() * 2 for t in (host["latents"], host["y"], host["clip"], host["audio"])) + \
        sum(t.numel() * 2 for t in host["ctx"])
    out_host = torch.empty_like(host["latents"]).pin_memory()

    def to_dev():
        return dict(latents=host["latents"].to(dev, non_blocking=True), y=host["y"].to(dev, non_blocking=True),
                    clip=host["clip"].to(dev, non_blocking=True), audio=host["audio"].to(dev, non_blocking=True),
                    ctx=[c.to(dev, non_blocking=True) for c in host["ctx"]])

    def step(d, i):
        t = sched._timesteps_host[i % 50]
        return pipe.denoise_step(d["latents"], t, sched.dsigma_at(i % 50), d["ctx"], d["clip"], d["y"], d["audio"], seq_len=L,
                                 clip_length=args.frames, text_guide_scale=3.0, audio_guide_scale=5.0, do_cfg=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = to_dev()

    # ---- multi-GPU parity, driver-visible: the sequence-parallel forward against the single-GPU forward of the same
    # inputs on every rank (the single-GPU semantics are the oracle of the SP path), max over ranks
    sp_parity = None
    if world > 1:
        from stableavatar_b200.pipeline import _frames_kwarg
        kw = dict(x=resident["latents"][:1].expand(3, -1, -1, -1, -1).contiguous(), t=torch.full((3,), 900.0, device=dev),
                  context=resident["ctx"], seq_len=L, clip_fea=resident["clip"], y=resident["y"],
                  vocal_embeddings=resident["audio"][:3], **_frames_kwarg(model, args.frames))
        single = model(**kw).float()
        model.enable_multi_gpus_inference()
        sp_out = model(**kw).float()
        torch.cuda.synchronize()
        err = torch.tensor([rel_l2(sp_out, single)], device=dev)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        sp_parity = err.item()
        del single, sp_out, kw
        if not (sp_parity <= 1e-2):
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "sequence-parallel forward differs from the single-GPU forward",
                                  "sp_parity_rel_l2": sp_parity, "n_gpus": world}), flush=True)
            dist.destroy_process_group()
            raise SystemExit(3)

    for i in range(args.warmup):
        step(resident, i)
    barrier()

    # ---- timed region 1: inputs resident in HBM ("value"): K steps of the pipeline's default path (one captured CUDA
    # graph replayed per step; eager launches with --no-graph)
    pipe.use_cuda_graphs = not args.no_graph
    for i in range(2):
        step(resident, i)                                    # graph capture + one replay, untimed
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # (read on rank 0 only and BEFORE the barrier: nvidia-smi takes ~1 s, and a rank that starts its timed steps late makes
    # every other rank wait in the first exchange — the 2-GPU run of this round measured 0.855 instead of 0.404 s/step so)
    nvl0 = nvlink_counters(local) if (world > 1 and rank == 0) else None
    barrier()
    e0.record()
    for i in range(args.steps):
        step(resident, i)
    e1.record()
    barrier()
    nvl1 = nvlink_counters(local) if nvl0 is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    clocks = sampler.summary() if sampler else None

    # ---- instrumented pass: the same K steps launched eagerly, with CUDA events around the dominant kernels on the
    # launching stream (events cannot be read back from inside a replayed graph) and the launch counter running.
    # Under sequence parallelism a second pass times the optional per-CFG-sample exchange pipeline (model.sp_pipelined).
    def instrumented():
        ops.TIMING = {}
        L_.launch_count = 0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for i in range(args.steps):
            step(resident, i)
        b.record()
        barrier()
        timing, ops.TIMING = ops.TIMING, None
        total = a.elapsed_time(b)
        return total, {k: sum(x.elapsed_time(y) for x, y in v) for k, v in timing.items()}, \
            {k: len(v) for k, v in timing.items()}, L_.launch_count

    pipe.use_cuda_graphs = False
    ms_eager, tag_ms, tag_n, launches = instrumented()
    serial = None
    if world > 1:                 # second eager pass with the per-sample exchange pipeline (the shipped default is serial)
        serial = dict(total=ms_eager, ms=tag_ms, n=tag_n)
        model.sp_pipelined = True
        ms_eager_pp, pp_ms, _, _ = instrumented()
        model.sp_pipelined = False

    # ---- timed region 2: end to end through the pipeline API with host buffers ("e2e"): every step copies its inputs
    # from pinned host memory, runs the captured step (the conditioning changed, so the context is re-encoded) and reads
    # the result back
    pipe.use_cuda_graphs = not args.no_graph
    staged = to_dev()
    for i in range(2):
        step(staged, i)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        fresh = to_dev()                                     # this step's inputs from pinned host memory
        for k in ("latents", "y", "clip", "audio"):
            staged[k].copy_(fresh[k])
        for a, bb in zip(staged["ctx"], fresh["ctx"]):
            a.copy_(bb)
        out_host.copy_(step(staged, i), non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the caller reads the step's result on the host
    e3.record()
    barrier()
    ms_e2e = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    s_per_step = ms.item() / 1e3 / args.steps
    s_e2e = ms_e2e.item() / 1e3 / args.steps

    extras = {"vae_decode_s": None, "vae_pp_equal": None, "clip_s": None}

    def vae_stage():
        """Config 4: one Wan VAE decode of the final latents (pipe.py:793-799). One GPU: plain decode. N > 1: the decoder
        runs as a pipeline over the N ranks (AutoencoderKLWan.enable_multi_gpus_decode) and its frames are compared
        bit for bit with the single-GPU decode of the same latents on every rank; time = max over ranks."""
        from stableavatar_b200.wan_vae import AutoencoderKLWan
        vae = AutoencoderKLWan()
        vae.load_state_dict(synth.vae_state_dict(encoder=True), strict=True)
        vae = vae.to(dev)
        z = synth.det_normal("bench_z", (1, 16, F_lat, h, w)).to(dev)
        single = None
        if world > 1:
            single = vae.decode(z).sample
            vae.enable_multi_gpus_decode()
        vae.decode(z[:, :, :2])                              # warm-up: operand preparation
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        video = vae.decode(z).sample
        ev1.record()
        barrier()
        assert video.shape == (1, 3, args.frames, args.height, args.width)
        tv = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            same = torch.tensor([int(torch.equal(video, single))], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            extras["vae_pp_equal"] = bool(same.item())
        extras["vae_decode_s"] = tv.item() / 1e3
        del video, single
        return vae

    def clip_stage(vae):
        """BASELINE "e2e s/clip", measured: ONE call of the reference entry point `pipe(...)` with pre-encoded prompt /
        CLIP / audio features (the encoders are outside the path): VAE encode of the conditioning clip, mask / y
        assembly, 50 denoise steps with 3-way CFG, VAE decode, the 388 MB fp32 read-back of decode_latents. Wall clock
        between synchronised barriers, max over ranks."""
        pipe.vae = vae
        gen = torch.Generator(device=dev).manual_seed(0)
        pos, neg = host["ctx"][2].to(dev), host["ctx"][0].to(dev)
        audio = resident["audio"][1:2]
        cond = synth.det_normal("bench_cond_image", (1, 3, 1, args.height, args.width)).clamp_(-1, 1)
        kw = dict(height=args.height, width=args.width, num_frames=args.frames, clip_length=args.frames, guidance_scale=6.0,
                  text_guide_scale=3.0, audio_guide_scale=5.0, generator=gen, prompt_embeds=[pos], negative_prompt_embeds=[neg],
                  clip_context=resident["clip"][:1], cond_image=cond, vocal_input_values=torch.zeros(args.frames * 640),
                  sr=16000, fps=25, vocal_embeddings_fn=lambda ws, we, last: audio, overlap_window_length=5)
        pipe(num_inference_steps=1, **kw)                    # warm-up: VAE encode operand preparation
        barrier()
        t0 = time.perf_counter()
        video = pipe(num_inference_steps=args.clip_steps, **kw).videos
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert tuple(video.shape) == (1, 3, args.frames, args.height, args.width) and bool(torch.isfinite(video).all())
        extras["clip_s"] = dt.item()

    def emit():
        if rank == 0:
            c = line["config"]
            c["vae_decode_s"], c["clip_s"] = extras["vae_decode_s"], extras["clip_s"]
            c["clip_s_model"] = None if extras["vae_decode_s"] is None else 50 * s_per_step + extras["vae_decode_s"]
            if world > 1:
                line["vae_pp_equal"] = extras["vae_pp_equal"]
            print(json.dumps(line), flush=True)

    if rank == 0:
        peaks = load_peaks()
        total_flops, attn_flops_per_launch = step_flops(cfg, L)
        total_flops, attn_flops_per_launch = total_flops * Wn, attn_flops_per_launch * Wn
        src = serial if serial is not None else dict(total=ms_eager, ms=tag_ms, n=tag_n)
        n_attn = src["n"].get("self_attn", 0)
        attn_avg = src["ms"].get("self_attn", 0.0) / max(1, n_attn)
        achieved = attn_flops_per_launch / world / (attn_avg * 1e-3) / 1e12 if n_attn else None
        shares = {k: v / ms_eager for k, v in tag_ms.items()}
        shares["other"] = max(0.0, 1.0 - sum(shares.values()))
        line = {
            "metric": metric_name(args), "value": s_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args, L),
                       "detail": f"{cfg['num_layers']} blocks, CFG batch 3 + CFG/Euler, text 512 + CLIP 257 + audio 21x15; text/CLIP context encoded "
                                 "once per clip (value) / once per step (e2e: its inputs change every step)",
                       "parallelism": f"sp{world}" if world > 1 else "single", "windows_per_step": Wn,
                       "l2_policy": "activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                       "step_tflop": total_flops / 1e12,
                       "step_tflops_achieved": total_flops / s_per_step / 1e12 / world,
                       "bf16_peak_frac_step": total_flops / s_per_step / 1e12 / world / peaks["bf16"],
                       "launch_mode": "eager" if args.no_graph else "cuda-graph replay (value, e2e; under sequence parallelism the self-attention "
                                      "exchange is peer-store kernels on side streams inside the graph, one eager NCCL all-gather per step); "
                                      "kernel timing and gpu_launches from an eager pass of the same steps",
                       "eager_ms_per_step": ms_eager / args.steps, "kernel_time_share": shares,
                       "vae_decode_s": None, "clip_s": None, "clip_s_model": None,
                       "clip_s_note": f"clip_s = one measured pipe(...) call: VAE encode of the conditioning clip + {args.clip_steps} denoise steps + "
                                      "VAE decode + fp32 read-back (wall clock, max over ranks); clip_s_model = 50 x value + vae_decode_s"
                                      + (f"; decoder pipelined over the {world} GPUs" if world > 1 else "")},
            "roofline": {"kernel": "attn8::flash_attn_v8_kernel (self-attention, sa_flash_attn_d128)", "bound": "tensor", "achieved": achieved,
                         "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"] if achieved else None,
                         "traffic": SELF_ATTN_DRAM_BYTES_B3 if (world == 1 and (args.frames, args.height, args.width) == (81, 480, 832)) else None,
                         "traffic_source": "profiles/r02_selfattn_in_bench_ncu.txt (ncu --set full, dram__bytes_read+write of one "
                                           "B=3 self-attention launch; algorithmic q+k+v+o = 1208 MB); null at N > 1 (ncu is single-GPU only "
                                           "and the 4x2 split at N = 8 reads each K/V chunk on two ranks)",
                         "peak_source": f"{peaks['src']} sustained bf16 (MEASURED_PEAKS.json)",
                         "launches_timed": n_attn, "avg_launch_ms": attn_avg,
                         "timed_in": "eager pass"},
            "e2e": {"value": s_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": out_host.numel() * 2},
            "gpu_launches": launches, "clocks": clocks,
        }
        if world > 1:
            line["sp_parity_rel_l2"] = sp_parity
            per_step = lambda d, k: d["ms"].get(k, 0.0) / args.steps  # noqa: E731
            region = pp_ms.get("sp_attn_region", 0.0) / args.steps
            attn = per_step(serial, "self_attn")
            exposed = per_step(serial, "sp_a2a_qkv") + per_step(serial, "sp_a2a_o")
            line["sp_exchange"] = {
                "note": "per step and rank, ms, eager passes. qkv / o: scatter + flag barrier around the attention kernel (the "
                        "shipped serial order); region_pipelined: the optional per-CFG-sample pipeline (model.sp_pipelined), "
                        "fork -> exchange || attention per sample -> join + barrier",
                "qkv_ms": per_step(serial, "sp_a2a_qkv"), "o_ms": per_step(serial, "sp_a2a_o"), "attention_ms": attn,
                "exposed_ms": exposed, "exposed_share_of_step": exposed / (ms_eager / args.steps),
                "region_pipelined_ms": region, "eager_ms_per_step_pipelined": ms_eager_pp / args.steps,
                "nvlink_tx_bytes_per_step_gpu0": None if not (nvl0 and nvl1) else (nvl1[0] - nvl0[0]) / args.steps,
                "nvlink_rx_bytes_per_step_gpu0": None if not (nvl0 and nvl1) else (nvl1[1] - nvl0[1]) / args.steps,
                "nvlink_note": "nvidia-smi nvlink -gt d on GPU 0 around the timed region (driver counters, KiB granularity); "
                               "algorithmic egress per step and rank = 30 blocks x (q + k,v to every rank of the head group + o)"}
    case = None
    if world == 1 and not args.no_cpu_baseline and args.model == "1.3b":
        # CPU baseline: one oracle block at the full sequence length on all host threads — and the same block on the B200
        # on the same numbers: parity at the benchmarked size
        cores = host_threads()
        case = cpu_block_case(cfg, L, (F_lat, h // 2, w // 2))
        per_block, ref_out = cpu_block_run(case)
        line["cpu_baseline"] = {"value": per_block * cfg["num_layers"] * 3, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"1 WanAttentionBlock of the oracle port (fp32) at L={L}, B=1 measured "
                                          f"{per_block:.2f} s on {cores} threads; extrapolated x{cfg['num_layers']} blocks x3 CFG samples"}
        gpu_out = gpu_block_parity(cfg, case, dev)
        line["block_parity_rel_l2"] = rel_l2(gpu_out, ref_out)
        line["block_parity_note"] = ("B200 WanAttentionBlock vs oracle block (fp32 CPU) on the same bf16-representable weights and "
                                     f"inputs at dim {cfg['dim']}, ffn {cfg['ffn_dim']}, L={L}, B=1; north_star bar 2e-2")
        del gpu_out, ref_out, case
    # The VAE and clip stages run last and under a watchdog: if one does not finish, the step line is still printed.
    if args.no_vae:
        emit()
    else:
        def give_up():
            emit()
            os._exit(0)
        timer = threading.Timer(args.stage_timeout, give_up)
        timer.daemon = True
        timer.start()
        try:
            vae = vae_stage()
            if not args.no_clip:
                clip_stage(vae)
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] VAE / clip stage failed on rank {rank}: {exc!r}", file=sys.stderr, flush=True)
        timer.cancel()
        emit()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and "block_parity_rel_l2" in line and not (line["block_parity_rel_l2"] <= 2e-2):
        raise SystemExit(4)
    if world > 1 and extras["vae_pp_equal"] is False:
        raise SystemExit(5)


def workload_name(args, L):
    name = "14B (train_14B architecture)" if args.model == "14b" else "1.3B"
    return f"{name} audio-DiT denoise step, {args.height}x{args.width}x{args.frames}f, CFG batch 3, L={L}"


def metric_name(args):
    if args.model == "14b":
        return f"s/denoise-step 14B @{args.height}x{args.width}x{args.frames}f"
    return METRIC


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="1.3b", choices=["1.3b", "14b"],
                    help="14b: BASELINE config 5 (train_14B architecture, default 720x1280x81; VAE / clip stages skipped)")
    ap.add_argument("--windows", type=int, default=1, help="sliding-window windows batched into one forward per step")
    ap.add_argument("--frames", type=int, default=81)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="e2e region without CUDA-graph replay")
    ap.add_argument("--no-vae", action="store_true", help="skip the VAE decode and measured-clip stages")
    ap.add_argument("--no-clip", action="store_true", help="skip the measured pipe(...) clip (config.clip_s)")
    ap.add_argument("--clip-steps", type=int, default=50, help="denoise steps of the measured clip")
    ap.add_argument("--stage-timeout", type=float, default=420.0, help="watchdog for the VAE + clip stages, seconds")
    args = ap.parse_args()
    if args.height is None:
        args.height = 720 if args.model == "14b" else 480
    if args.width is None:
        args.width = 1280 if args.model == "14b" else 832
    if args.model == "14b":
        args.no_vae = True
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference)")
        run_b200(args)


if __name__ == "__main__":
    main()
