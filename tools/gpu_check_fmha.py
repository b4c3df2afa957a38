"""FMHA parity (vs fp32 SDPA) over the test shapes + timings at the DiT-L/2 shapes.  One kernel per process:
LN3_FMHA_KERNEL=2|3, LN3_FMHA_ROTA=0|1, LN3_FMHA_POLY=0|2, LN3_FMHA_TAIL=0|1 (read once by the library)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from ln3diff_b200 import ops

dev = "cuda"


def rel(a, b):
    return ((a.float().cpu() - b).norm() / b.norm()).item()


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


tag = " ".join(f"{k[9:].lower()}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("LN3_FMHA_"))
worst = 0.0
shapes = [(2, 12, 768, 768), (2, 16, 768, 77), (1, 4, 200, 333), (2, 16, 768, 1024), (3, 16, 256, 256), (1, 2, 1, 1),
          (13, 16, 768, 768), (16, 16, 700, 77), (9, 16, 300, 130), (1, 1, 384, 96), (1, 1, 385, 97), (5, 3, 129, 191),
          (16, 16, 768, 768), (24, 16, 256, 256), (8, 16, 768, 768)]
for (B, H, Lq, Lkv) in shapes:
    g = torch.Generator().manual_seed(Lq * 7 + Lkv)
    D = H * 64
    qkv = torch.randn(B, max(Lq, Lkv), 3 * D, generator=g).bfloat16()
    q, k, v = qkv[:, :Lq, :D], qkv[:, :Lkv, D:2 * D], qkv[:, :Lkv, 2 * D:]
    dq = qkv.to(dev)
    out = ops.fmha(dq[:, :Lq, :D], dq[:, :Lkv, D:2 * D], dq[:, :Lkv, 2 * D:], H)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().reshape(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, Lq, D)
    r = rel(out, ref)
    worst = max(worst, r)
    if r > 6e-3 or r != r:
        print(f"[{tag}] FAIL shape {(B, H, Lq, Lkv)} rel {r:.3e}", flush=True)
# second K/V source (I23D: 768 latent + 256 DINO tokens; and a ragged first source)
two = [(2, 4, 200, 256, 77), (2, 16, 768, 768, 256)] + ([] if os.environ.get('LN3_FMHA_KERNEL') == '2' else [(1, 2, 100, 100, 50)])
for (B, H, Lq, L1, L2) in two:
    g = torch.Generator().manual_seed(17 + L1)
    D = H * 64
    q = torch.randn(B, Lq, D, generator=g).bfloat16()
    k1, v1 = torch.randn(B, L1, D, generator=g).bfloat16(), torch.randn(B, L1, D, generator=g).bfloat16()
    k2, v2 = torch.randn(B, L2, D, generator=g).bfloat16(), torch.randn(B, L2, D, generator=g).bfloat16()
    out = ops.fmha(q.to(dev), k1.to(dev), v1.to(dev), H, k2=k2.to(dev), v2=v2.to(dev))
    sp = lambda t_: t_.float().reshape(B, -1, H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q), sp(torch.cat([k1, k2], 1)), sp(torch.cat([v1, v2], 1)))
    r = rel(out, ref.transpose(1, 2).reshape(B, Lq, D))
    worst = max(worst, r)
    if r > 6e-3 or r != r:
        print(f"[{tag}] FAIL two-source shape {(B, H, Lq, L1, L2)} rel {r:.3e}", flush=True)
# large-magnitude scores exercise the lazy rescale
g = torch.Generator().manual_seed(5)
B, H, L = 2, 4, 768
qkv = (torch.randn(B, L, 3 * H * 64, generator=g) * 3.0).bfloat16()
dq = qkv.to(dev)
D = H * 64
out = ops.fmha(dq[:, :, :D], dq[:, :, D:2 * D], dq[:, :, 2 * D:], H)
qf, kf, vf = (t.float().reshape(B, -1, H, 64).transpose(1, 2) for t in (qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]))
ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, L, D)
r = rel(out, ref)
worst = max(worst, r)
if r > 8e-3 or r != r:
    print(f"[{tag}] FAIL peaked rel {r:.3e}", flush=True)

torch.manual_seed(0)
B, H, L = 16, 16, 768
qkv = (torch.randn(B, L, 3 * H * 64, device=dev) * 0.5).bfloat16()
q, k, v = qkv[:, :, :H * 64], qkv[:, :, H * 64:2 * H * 64], qkv[:, :, 2 * H * 64:]
us = timeit(lambda: ops.fmha(q, k, v, H))
qc = (torch.randn(B // 2, L, H * 64, device=dev) * 0.5).bfloat16()
kvc = (torch.randn(B // 2, 77, 2 * H * 64, device=dev) * 0.5).bfloat16()
us2 = timeit(lambda: ops.fmha(qc, kvc[:, :, :H * 64], kvc[:, :, H * 64:], H))
# decoder shapes: in-plane (3B, 256, 256) and global (B, 768, 768) at 8 latents
qd = (torch.randn(24, 256, 3 * H * 64, device=dev) * 0.5).bfloat16()
us3 = timeit(lambda: ops.fmha(qd[:, :, :H * 64], qd[:, :, H * 64:2 * H * 64], qd[:, :, 2 * H * 64:], H))
# I23D self-attention: 768 latent + 256 DINO tokens, 16 samples
qi = (torch.randn(16, 768, 3 * H * 64, device=dev) * 0.5).bfloat16()
kd = (torch.randn(16, 256, 2 * H * 64, device=dev) * 0.5).bfloat16()
us4 = timeit(lambda: ops.fmha(qi[:, :, :H * 64], qi[:, :, H * 64:2 * H * 64], qi[:, :, 2 * H * 64:], H,
                              k2=kd[:, :, :H * 64], v2=kd[:, :, H * 64:]))
qg = (torch.randn(8, 768, 3 * H * 64, device=dev) * 0.5).bfloat16()
us5 = timeit(lambda: ops.fmha(qg[:, :, :H * 64], qg[:, :, H * 64:2 * H * 64], qg[:, :, 2 * H * 64:], H))
print(f"[{tag}] i23d(16x16x768x(768+256)) {us4:.1f} us; global(8x16x768x768) {us5:.1f} us")
print(f"[{tag}] worst rel {worst:.2e}; self(16x16x768x768) {us:.1f} us ({4.0 * B * H * L * L * 64 / us / 1e6:.0f} TF/s); "
      f"cross(8x16x768x77) {us2:.1f} us; inplane(24x16x256x256) {us3:.1f} us", flush=True)
