"""FMHA timings (self 768x768, cross 768x77) at the DiT-L/2 B'=16 shape; LN3_FMHA_POLY selects the variant."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import ops

dev = "cuda"
torch.manual_seed(0)
B, H, L = 16, 16, 768


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


qkv = (torch.randn(B, L, 3 * H * 64, device=dev) * 0.5).bfloat16()
q, k, v = qkv[:, :, :H * 64], qkv[:, :, H * 64:2 * H * 64], qkv[:, :, 2 * H * 64:]
us = timeit(lambda: ops.fmha(q, k, v, H))
out = ops.fmha(q, k, v, H)
qf, kf, vf = (t.float().reshape(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, L, H * 64)
rel = ((out.float() - ref).norm() / ref.norm()).item()
qc = (torch.randn(B, L, H * 64, device=dev) * 0.5).bfloat16()
kvc = (torch.randn(B, 77, 2 * H * 64, device=dev) * 0.5).bfloat16()
us2 = timeit(lambda: ops.fmha(qc, kvc[:, :, :H * 64], kvc[:, :, H * 64:], H))
print(f"mma2={os.environ.get('LN3_FMHA_MMA2','1')} tail={os.environ.get('LN3_FMHA_TAIL','1')} ptmem={os.environ.get('LN3_FMHA_PTMEM','0')} split={os.environ.get('LN3_FMHA_SPLIT','0')} pp={os.environ.get('LN3_FMHA_PINGPONG','0')} poly={os.environ.get('LN3_FMHA_POLY', 'default')} self {us:.1f} us ({4.0 * B * H * L * L * 64 / us / 1e6:.0f} TF/s) "
      f"rel {rel:.2e}; cross {us2:.1f} us", flush=True)
