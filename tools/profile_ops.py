"""Run ONE op of the DiT-L/2 B'=16 step a few times (for `ncu --set full` captures of a single kernel):
    python tools/profile_ops.py {qkv|fc1_gelu|fc1_noact|fc2|proj|fmha_self|fmha_cross|norm_ln_resid} [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import ops
from ln3diff_b200._lib import NORM_LAYER

op = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda"
torch.manual_seed(0)
M, D = 12288, 1024
GEMMS = {"qkv": (3072, 1024, ops.ACT_NONE), "fc1_gelu": (4096, 1024, ops.ACT_GELU_ERF), "fc1_noact": (4096, 1024, ops.ACT_NONE),
         "fc2": (1024, 4096, ops.ACT_NONE), "proj": (1024, 1024, ops.ACT_NONE)}
if op in GEMMS:
    N, K, act = GEMMS[op]
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.03).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, b, act=act, out=out)
elif op.startswith("fmha"):
    B, H, L = 16, 16, 768
    if op == "fmha_self":
        qkv = (torch.randn(B, L, 3 * H * 64, device=dev) * 0.5).bfloat16()
        q, k, v = qkv[:, :, :H * 64], qkv[:, :, H * 64:2 * H * 64], qkv[:, :, 2 * H * 64:]
    else:
        q = (torch.randn(8, L, H * 64, device=dev) * 0.5).bfloat16()
        kv = (torch.randn(8, 77, 2 * H * 64, device=dev) * 0.5).bfloat16()
        k, v = kv[:, :, :H * 64], kv[:, :, H * 64:]
    fn = lambda: ops.fmha(q, k, v, H)
elif op == "norm_ln_resid":
    x = torch.randn(M, D, device=dev)
    val = torch.randn(M, D, device=dev).bfloat16()
    mod = torch.randn(16, 3 * D, device=dev)
    out = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.norm_modulate(x, norm=NORM_LAYER, shift=mod[:, :D], scale=mod[:, D:2 * D], mod_rows=768, out=out,
                                   resid=val, resid_gate=mod[:, 2 * D:], resid_gate_rows=768)
else:
    raise SystemExit(f"unknown op {op}")
for _ in range(n):
    fn()
torch.cuda.synchronize()
print("done")
