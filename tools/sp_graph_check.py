"""2-GPU check: segmented CUDA-graph capture + replay of the sequence-parallel DiT forward (NCCL calls eager between segments).
Run: timeout 150 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/sp_graph_check.py"""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from stableavatar_b200 import synth
from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
cfg = synth.DIT_TINY
keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim", "num_heads", "num_layers")
m = WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys})
m.load_state_dict({k: v.bfloat16() for k, v in synth.dit_state_dict(cfg).items()}, strict=True)
m = m.to(dev, torch.bfloat16)
inp = synth.dit_inputs(cfg, frames=17, height=128, width=192, seed=5)
bf = torch.bfloat16
kw = dict(x=inp["x"].to(dev, bf), t=inp["t"].to(dev), context=[c.to(dev, bf) for c in inp["context"]], seq_len=inp["seq_len"],
          clip_fea=inp["clip_fea"].to(dev, bf), y=inp["y"].to(dev, bf), vocal_embeddings=inp["vocal_embeddings"].to(dev, bf),
          video_sample_n_frames=17)
single = m(**kw).float()
m.enable_multi_gpus_inference()
eager = m(**kw).float()
torch.cuda.synchronize()
print(rank, "eager sp vs single", ((eager - single).norm() / single.norm()).item(), flush=True)
from stableavatar_b200.sequence_parallel import SegmentedGraph
print(rank, "capturing segmented graph", flush=True)
sg = SegmentedGraph(lambda: m(**kw), device=dev)
print(rank, "captured", len(sg.segments), "segments", flush=True)
for i in range(3):
    out = sg.replay()
    torch.cuda.synchronize()
    print(rank, "replay", i, ((out.float() - single).norm() / single.norm()).item(), flush=True)
dist.barrier()
dist.destroy_process_group()
