"""Run a few big GEMMs (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import ops

M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (12288, 3072, 1024)))
dev = "cuda"
a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
b = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
act = int(sys.argv[4]) if len(sys.argv) > 4 else ops.ACT_NONE
for _ in range(4):
    ops.gemm(a, w, b, out=out, act=act)
torch.matmul(a, w.t())
torch.cuda.synchronize()
print("done")
