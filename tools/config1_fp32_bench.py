"""BASELINE config 1 on the GPU in fp32 mode: the full 1.3B audio-DiT (30 layers, random-init float32 weights), one
denoise evaluation at 480x832x5 frames (L = 3120), CFG batch 3 — the case the reference runs on the CPU."""
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import step_flops  # noqa: E402
from stableavatar_b200 import _lib, synth  # noqa: E402
from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel  # noqa: E402

cfg = synth.DIT_1_3B
keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim", "num_heads",
        "num_layers")
with torch.device("cuda"):
    model = WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys})
model.init_random_(seed=0)
assert model.dtype == torch.float32
inp = synth.dit_inputs(cfg, frames=5, height=480, width=832)
args = dict(x=inp["x"].cuda(), t=inp["t"].cuda(), context=[c.cuda() for c in inp["context"]], seq_len=inp["seq_len"],
            clip_fea=inp["clip_fea"].cuda(), y=inp["y"].cuda(), vocal_embeddings=inp["vocal_embeddings"].cuda(),
            video_sample_n_frames=5)
out = model(**args)                       # warm-up: weight splits are built and cached here
torch.cuda.synchronize()
_lib.launch_count = 0
t0 = time.perf_counter()
out = model(**args)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
flop, _ = step_flops(cfg, inp["seq_len"], G=2, A=15)
print(f"config 1 (fp32 mode, B=3, L={inp['seq_len']}): {dt:.3f} s per evaluation, {flop / 1e12:.1f} TFLOP algorithmic "
      f"({6 * flop / dt / 1e12:.0f} TFLOP/s of bf16 tensor work incl. the 6x split), {_lib.launch_count} launches, "
      f"finite={bool(torch.isfinite(out).all())}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
