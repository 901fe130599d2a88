"""Full-size Wan VAE decode timing on one B200: z [1,16,T,60,104] -> [1,3,1+4(T-1),480,832] (BASELINE config 4)."""
import sys
import time
import torch
sys.path.insert(0, ".")
from stableavatar_b200 import synth, _lib
from stableavatar_b200 import wan_vae
from stableavatar_b200.wan_vae import AutoencoderKLWan

GRAPH = "--graph" in sys.argv            # opt-in CUDA-graph replay of the steady-state chunk (the second decode captures)
ENCODE = "--encode" in sys.argv
sys.argv = [a for a in sys.argv if a not in ("--graph", "--encode")]
if "--no-halo" in sys.argv:            # A/B: every conv on the per-tap kernel
    wan_vae._Conv.use_halo = False
    sys.argv.remove("--no-halo")

T = int(sys.argv[1]) if len(sys.argv) > 1 else 21
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (60, 104)
vae = AutoencoderKLWan()
vae.load_state_dict(synth.vae_state_dict(), strict=True)
vae = vae.to("cuda")
vae.use_cuda_graph = GRAPH
z = synth.det_normal("vae_zfull", (1, 16, T, h, w)).cuda()
out = vae.decode(z).sample          # warm-up (allocations, attribute setup)
torch.cuda.synchronize()
_lib.launch_count = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
out = vae.decode(z).sample
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
flop = (4.29 + (T - 1) * 13.49 + T * 0.06) * 1e12 * (h * w) / (60 * 104)
print(f"vae decode T={T} {h}x{w}: {ms:.1f} ms (wall {1e3 * (time.perf_counter() - t0):.1f} ms), {flop / ms / 1e9:.1f} TFLOP/s, "
      f"{_lib.launch_count} kernel launches, out {tuple(out.shape)} finite={bool(torch.isfinite(out).all())} "
      f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

if ENCODE:
    # Encode of the conditioning clip: x [1,3,1+4(T-1),8h,8w] -> latent [1,16,T,h,w]; 162.5 TFLOP at
    # the full size (81 x 480 x 832).
    vae.load_state_dict(synth.vae_state_dict(encoder=True), strict=True)
    F = 1 + 4 * (T - 1)
    x = out.clamp(-1, 1) if out.shape[2] == F else torch.zeros(1, 3, F, 8 * h, 8 * w, device="cuda")
    del out
    lat = vae.encode(x).latent_dist.mode()
    torch.cuda.synchronize()
    _lib.launch_count = 0
    e0.record()
    lat = vae.encode(x).latent_dist.mode()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flop = 162.5e12 * (F / 81) * (h * w) / (60 * 104)
    print(f"vae encode F={F} {8 * h}x{8 * w}: {ms:.1f} ms, {flop / ms / 1e9:.1f} TFLOP/s, {_lib.launch_count} kernel launches, "
          f"latent {tuple(lat.shape)} finite={bool(torch.isfinite(lat).all())} mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
