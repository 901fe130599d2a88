"""Pipeline-parallel VAE decode timing (torchrun, one rank per GPU): full clip 21x60x104 -> 81x480x832."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from stableavatar_b200 import synth
from stableavatar_b200.wan_vae import AutoencoderKLWan
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
vae = AutoencoderKLWan()
vae.load_state_dict(synth.vae_state_dict(), strict=True)
vae = vae.to(dev)
vae.enable_multi_gpus_decode()
z = synth.det_normal("vae_zfull", (1, 16, 21, 60, 104)).to(dev)
vae.decode(z[:, :, :3])
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = vae.decode(z).sample
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"pipeline-parallel VAE decode on {world} GPUs: {ms.item():.1f} ms (max over ranks), finite={bool(torch.isfinite(out).all())}, ranges={vae.partition_units(vae._unit_costs(60, 104), world)}")
dist.destroy_process_group()
