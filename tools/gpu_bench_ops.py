"""Per-op timings of the DiT-L/2 B'=16 forward shapes (CUDA events around 20 back-to-back launches).

Run on the B200 box:  python tools/gpu_bench_ops.py  -> gpurun_out/bench_ops.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import ops
from ln3diff_b200._lib import NORM_LAYER, NORM_NONE

dev = "cuda"
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    """Back-to-back launches inside one event pair (launch latency hidden; caches warm)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


res = {}
M, D = 12288, 1024
x32 = torch.randn(M, D, device=dev)
xb = x32.bfloat16()


def gemm_case(name, N, K, act=ops.ACT_NONE, M=M):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.03).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: ops.gemm(a, w, b, act=act, out=out))
    res[name] = {"us": us, "tflops": 2.0 * M * N * K / us / 1e6}


gemm_case("gemm_qkv_3072x1024", 3072, 1024)
gemm_case("gemm_fc1_gelu_4096x1024", 4096, 1024, ops.ACT_GELU_ERF)
gemm_case("gemm_fc1_noact_4096x1024", 4096, 1024)
gemm_case("gemm_fc2_1024x4096", 1024, 4096)
gemm_case("gemm_proj_1024x1024", 1024, 1024)
gemm_case("gemm_q_cond_1024x1024_m6144", 1024, 1024, M=6144)

B, H, L = 16, 16, 768
qkv = (torch.randn(B, L, 3 * H * 64, device=dev) * 0.5).bfloat16()
q, k, v = qkv[:, :, :H * 64], qkv[:, :, H * 64:2 * H * 64], qkv[:, :, 2 * H * 64:]
us = timeit(lambda: ops.fmha(q, k, v, H))
res["fmha_self_768"] = {"us": us, "tflops": 4.0 * B * H * L * L * 64 / us / 1e6}
qc = torch.randn(B, L, H * 64, device=dev).bfloat16()
kvc = torch.randn(B, 77, 2 * H * 64, device=dev).bfloat16()
us = timeit(lambda: ops.fmha(qc, kvc[:, :, :H * 64], kvc[:, :, H * 64:], H))
res["fmha_cross_77"] = {"us": us, "tflops": 4.0 * B * H * L * 77 * 64 / us / 1e6}

mod = torch.randn(16, 6 * D, device=dev)
gate = (mod[:, :D] * 0.0).contiguous()
# rotate over 5 buffer sets (500 MB) so that every call streams from HBM as it does inside a forward
sets = [(torch.randn(M, D, device=dev), torch.randn(M, D, device=dev).bfloat16(),
         torch.empty(M, D, device=dev, dtype=torch.bfloat16)) for _ in range(5)]
cnt = [0]


def nm(kind):
    xr, val, out = sets[cnt[0] % 5]
    cnt[0] += 1
    if kind == "ln_resid":
        ops.norm_modulate(xr, shift=mod[:, D:2 * D], scale=mod[:, 2 * D:3 * D], mod_rows=768, norm=NORM_LAYER,
                          resid=val, resid_gate=gate, resid_gate_rows=768, out=out)
    elif kind == "cast_resid":
        ops.norm_modulate(xr, norm=NORM_NONE, resid=val, resid_gate=gate, resid_gate_rows=768, out=out)
    else:
        ops.norm_modulate(xr, shift=mod[:, D:2 * D], scale=mod[:, 2 * D:3 * D], mod_rows=768, norm=NORM_LAYER, out=out)


for kind, bytes_per in (("ln_resid", 12), ("cast_resid", 12), ("ln", 6)):
    us = timeit(lambda: nm(kind), iters=40)
    res["norm_modulate_" + kind] = {"us": us, "gbps": M * D * bytes_per / us / 1e3}

xin = torch.randn(16, 12, 32, 32, device=dev)
w = torch.randn(D, 4, 2, 2, device=dev).contiguous()
bb = torch.randn(D, device=dev)
pe = torch.randn(768, D, device=dev)
us = timeit(lambda: ops.patch_embed(xin, w, bb, pe))
res["patch_embed"] = {"us": us}

for k_, v_ in res.items():
    print(k_, {a: round(b, 1) for a, b in v_.items()}, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_ops.json", "w"), indent=1)
