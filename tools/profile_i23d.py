"""Run N forward_with_cfg evaluations of the I23D DiT-PixArt-L/2 at the configs[3] shard (8 images -> 16 samples), eager
launches (LN3_CUDA_GRAPH=0) so that an ncu launch list sees the kernels:  python tools/profile_i23d.py [n]"""
import os
import sys

os.environ.setdefault("LN3_CUDA_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200.utils import build_i23d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda"
m = build_i23d("DiT-PixArt-L/2", device=dev)
g = torch.Generator(device=dev).manual_seed(0)
B = 8
z = torch.randn(B, 12, 32, 32, device=dev, generator=g)
c = {"vector": torch.randn(B, 768, device=dev, generator=g), "crossattn": torch.randn(B, 256, 2048, device=dev, generator=g)}
ctx = {k: torch.cat([v, torch.zeros_like(v)]) for k, v in c.items()}
x = torch.cat([z, z])
t = torch.full((2 * B,), 0.3, device=dev)
for _ in range(n):
    out = m.forward_with_cfg(x, t, ctx, 4.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    out = m.forward_with_cfg(x, t, ctx, 4.0)
e1.record()
torch.cuda.synchronize()
print("done", tuple(out.shape), f"{e0.elapsed_time(e1) / n:.3f} ms per forward_with_cfg (16 samples), graph={os.environ['LN3_CUDA_GRAPH']}, checksum {float(out.double().abs().mean()):.6f}")
