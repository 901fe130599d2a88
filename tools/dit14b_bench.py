"""One denoise evaluation of the train_14B architecture at BASELINE config 5 (dim 5120 / 40 heads / 40 layers / ffn 13824,
720x1280x81 frames -> latent 21x90x160, L = 75 600, CFG batch 3) on ONE B200, random-init bf16 weights (38 GB).
Usage: python tools/dit14b_bench.py [layers] [height width]   (fewer layers = proportional extrapolation, stated)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import step_flops  # noqa: E402
from stableavatar_b200 import _lib, ops, synth  # noqa: E402
from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasy14BModel  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 40
height, width = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (720, 1280)
cfg = dict(synth.DIT_14B, num_layers=layers)
keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim", "num_heads",
        "num_layers")
dev, bf = "cuda", torch.bfloat16
torch.set_default_dtype(bf)
with torch.device(dev):
    model = WanTransformer3DFantasy14BModel(**{k: cfg[k] for k in keys})
torch.set_default_dtype(torch.float32)
model.init_random_(seed=0)
inp = synth.dit_inputs(cfg, frames=81, height=height, width=width)
L = inp["seq_len"]
args = dict(x=inp["x"].to(dev, bf), t=inp["t"].to(dev), context=[c.to(dev, bf) for c in inp["context"]], seq_len=L,
            clip_fea=inp["clip_fea"].to(dev, bf), y=inp["y"].to(dev, bf), vocal_embeddings=inp["vocal_embeddings"].to(dev, bf))
print(f"14B x{layers} layers: {sum(p.numel() for p in model.parameters()) / 1e9:.2f} B params, L = {L}, "
      f"mem after init {torch.cuda.memory_allocated() / 2**30:.1f} GiB", flush=True)
out = model(**args)                      # warm-up: operand preparation, lazy tables
torch.cuda.synchronize()
_lib.launch_count = 0
ops.TIMING = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
out = model(**args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
flop, attn_flop = step_flops(cfg, L)
# the 14B class runs the adapter on all three samples (no [0, vc, vc] replication): 3x the adapter term of step_flops
adapter = 2 * (2 * 2 * L * cfg["dim"] ** 2) + 2 * 4 * 21 * 15 * (L // 21) * cfg["dim"]
flop += 2 * adapter
print(f"forward (B=3): {ms / 1e3:.3f} s, {flop / 1e12:.0f} TFLOP -> {flop / ms / 1e9:.0f} TFLOP/s "
      f"({flop / ms / 1e9 / 1391.5:.2f} of sustained bf16 peak), {_lib.launch_count} launches, "
      f"finite={bool(torch.isfinite(out.float()).all())}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, "
      f"wall {time.perf_counter() - t0:.2f} s", flush=True)
for tag, evs in (ops.TIMING or {}).items():
    tot = sum(a.elapsed_time(b) for a, b in evs)
    print(f"  {tag}: {tot:.1f} ms over {len(evs)} launches ({tot / ms:.1%} of the forward)")
if layers != 40:
    print(f"extrapolated to 40 layers: {ms / 1e3 * 40 / layers:.2f} s/step (embeddings/adapter counted {40 / layers:.0f}x: upper bound)")
