"""Repeated full-size VAE decodes in one process: CUDA-graph replay on/off x halo-staged conv on/off, with clocks / power after each group."""
import sys, subprocess, torch
sys.path.insert(0, ".")
from stableavatar_b200 import synth, wan_vae
from stableavatar_b200.wan_vae import AutoencoderKLWan
vae = AutoencoderKLWan(); vae.load_state_dict(synth.vae_state_dict(), strict=True); vae = vae.to("cuda")
z = synth.det_normal("vae_zfull", (1, 16, 21, 60, 104)).cuda()
for mode in ("graph+halo", "nograph+halo", "graph+halo", "nograph+nohalo"):
    vae.use_cuda_graph = mode.startswith("graph")
    wan_vae._Conv.use_halo = "+halo" in mode
    vae._prep = None
    vae.decode(z); vae.decode(z); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); vae.decode(z); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print(mode, " ".join(f"{t:.0f}" for t in ts), "|", clk, flush=True)
