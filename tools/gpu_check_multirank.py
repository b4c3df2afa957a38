"""N-rank plumbing check (torchrun): NCCL init, per-rank prompt shard through the sampling pipeline on a small
DiT, one all-gather of the finished latents, max-over-ranks timing -- the multi-GPU skeleton of bench.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ln3diff_b200 import pipeline
from ln3diff_b200.utils import build_t23d

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
B = 2
m = build_t23d("DiT-B/2", device=dev)
g = torch.Generator().manual_seed(0)
noise = torch.randn(world * B, 12, 32, 32, generator=g)              # one global draw, sliced per rank
c = torch.randn(world * B, 77, 768, generator=g)
sl = slice(rank * B, (rank + 1) * B)
z = noise[sl].to(dev)
cc, uc = {"crossattn": c[sl].to(dev)}, {"crossattn": torch.zeros(B, 77, 768, device=dev)}
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lat = pipeline.sample_t23d(m, z, cc, uc, 12, 6.5)
out = [torch.empty_like(lat) for _ in range(world)]
dist.all_gather(out, lat)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
allv = torch.cat(out)
ok = bool(torch.isfinite(allv).all()) and allv.shape[0] == world * B
# rank r's slice of the gathered tensor must equal what rank r computed
ok = ok and torch.equal(allv[sl], lat)
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} ok={bool(flag.item())} max_ms={ms.item():.1f} gathered={tuple(allv.shape)}", flush=True)
dist.destroy_process_group()
