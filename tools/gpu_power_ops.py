"""Sustained (power-capped) cost of the hot ops: each op runs back to back for ~1.5 s while NVML is sampled, so the
numbers are at the clocks the power governor settles on -- the regime the 250-step sampler actually runs in --
not at the burst clocks a 20-launch micro-benchmark sees.  Reports us per call, mean SM clock, mean power and
energy per call (J = W x s); the step's energy budget is the sum over its launches.

    python tools/gpu_power_ops.py  -> gpurun_out/power_ops.json
"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import ops
from ln3diff_b200._lib import NORM_LAYER, NORM_NONE

import pynvml

pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda"
torch.manual_seed(0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.p, self.c = [], []

    def run(self):
        while not self.stop:
            try:
                self.p.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1e3)
                self.c.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
            except Exception:
                pass
            time.sleep(0.02)


def sustained(fn, seconds=1.5, chunk=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s = Sampler()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.start()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(chunk):
            fn()
        n += chunk
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    us = e0.elapsed_time(e1) * 1e3 / n
    k = len(s.p) // 3          # drop the ramp
    pw = sum(s.p[k:]) / max(1, len(s.p[k:]))
    ck = sum(s.c[k:]) / max(1, len(s.c[k:]))
    return {"us": us, "sm_mhz": ck, "watts": pw, "mJ_per_call": pw * us * 1e-3, "calls": n}


res = {}
idle = []
for _ in range(20):
    idle.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1e3)
    time.sleep(0.02)
res["idle_watts"] = sum(idle) / len(idle)
M, D = 12288, 1024


def gemm_case(name, N, K, act=ops.ACT_NONE, m=M):
    a = (torch.randn(m, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.03).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.empty(m, N, device=dev, dtype=torch.bfloat16)
    r = sustained(lambda: ops.gemm(a, w, b, act=act, out=out))
    r["tflops"] = 2.0 * m * N * K / r["us"] / 1e6
    r["pJ_per_flop"] = r["mJ_per_call"] * 1e9 / (2.0 * m * N * K)
    res[name] = r
    print(name, r, flush=True)


gemm_case("gemm_qkv_3072x1024", 3072, 1024)
gemm_case("gemm_fc1_gelu_4096x1024", 4096, 1024, ops.ACT_GELU_ERF)
gemm_case("gemm_fc2_1024x4096", 1024, 4096)
gemm_case("gemm_proj_1024x1024", 1024, 1024)
gemm_case("gemm_q_cond_1024x1024_m6144", 1024, 1024, m=6144)

B, Hh, L = 16, 16, 768
qkv = (torch.randn(B, L, 3 * Hh * 64, device=dev) * 0.5).bfloat16()
q, k, v = qkv[:, :, :Hh * 64], qkv[:, :, Hh * 64:2 * Hh * 64], qkv[:, :, 2 * Hh * 64:]
r = sustained(lambda: ops.fmha(q, k, v, Hh))
r["tflops"] = 4.0 * B * Hh * L * L * 64 / r["us"] / 1e6
res["fmha_self_768"] = r
print("fmha_self_768", r, flush=True)
qc = torch.randn(B // 2, L, Hh * 64, device=dev).bfloat16()
kvc = torch.randn(B // 2, 77, 2 * Hh * 64, device=dev).bfloat16()
r = sustained(lambda: ops.fmha(qc, kvc[:, :, :Hh * 64], kvc[:, :, Hh * 64:], Hh))
res["fmha_cross_77_b8"] = r
print("fmha_cross_77_b8", r, flush=True)

mod = torch.randn(16, 6 * D, device=dev)
gate = (mod[:, :D] * 0.0).contiguous()
sets = [(torch.randn(M, D, device=dev), torch.randn(M, D, device=dev).bfloat16(),
         torch.empty(M, D, device=dev, dtype=torch.bfloat16)) for _ in range(5)]
cnt = [0]


def nm(kind):
    xr, val, out = sets[cnt[0] % 5]
    cnt[0] += 1
    if kind == "ln_resid":
        ops.norm_modulate(xr, shift=mod[:, D:2 * D], scale=mod[:, 2 * D:3 * D], mod_rows=768, norm=NORM_LAYER,
                          resid=val, resid_gate=gate, resid_gate_rows=768, out=out)
    else:
        ops.norm_modulate(xr, norm=NORM_NONE, resid=val, resid_gate=gate, resid_gate_rows=768, out=out)


for kind in ("ln_resid", "cast_resid"):
    r = sustained(lambda: nm(kind))
    r["gbps"] = 12.0 * M * D / r["us"] / 1e3
    res["norm_modulate_" + kind] = r
    print("norm_modulate_" + kind, r, flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/power_ops.json", "w"), indent=1)
