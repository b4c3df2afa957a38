"""GPU diagnostic for ln3_gemm_bf16: correctness on a ladder of shapes + timing.

Run on the B200 box:  python tools/gpu_check_gemm.py  (writes gpurun_out/gemm_check.json)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ln3diff_b200 import ops

torch.manual_seed(0)
dev = "cuda"
res = {"device": torch.cuda.get_device_name(0), "cases": []}


def run_case(M, N, K, act=ops.ACT_NONE, out_kind=ops.OUT_BF16, bias=True, gate=False, out2=False):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    if act == ops.ACT_GELU_ERF:
        ref = torch.nn.functional.gelu(ref)
    elif act == ops.ACT_GELU_TANH:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    elif act == ops.ACT_SILU:
        ref = torch.nn.functional.silu(ref)
    kw = {}
    if out_kind == ops.OUT_RESID_F32:
        x0 = torch.randn(M, N, device=dev)
        x = x0.clone()
        kw["out"] = x
        if gate:
            rows = 64
            g = torch.randn((M + rows - 1) // rows, N, device=dev)
            kw["gate"], kw["gate_rows"] = g, rows
            ref = x0 + g.repeat_interleave(rows, 0)[:M] * ref
        else:
            ref = x0 + ref
        if out2:
            kw["out2"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out = ops.gemm(a, w, b, act=act, out_kind=out_kind, **kw)
    torch.cuda.synchronize()
    o = out.float()
    err = (o - ref).abs()
    rel = (o - ref).norm() / ref.norm()
    case = {"M": M, "N": N, "K": K, "act": act, "out_kind": out_kind, "max_abs": err.max().item(),
            "rel_l2": rel.item(), "ref_absmax": ref.abs().max().item()}
    if out2:
        case["out2_rel"] = ((kw["out2"].float() - ref).norm() / ref.norm()).item()
    if rel.item() > 2e-2 or not torch.isfinite(o).all():
        # where are the errors?  per 32-column chunk and per 8-row group
        bad = (err > 0.05 * ref.abs().max()).float()
        case["bad_frac"] = bad.mean().item()
        case["bad_rows_first"] = bad.sum(1).nonzero().flatten()[:16].tolist()
        case["bad_cols_first"] = bad.sum(0).nonzero().flatten()[:16].tolist()
        case["sample_out"] = o[:4, :8].tolist()
        case["sample_ref"] = ref[:4, :8].tolist()
    print(case, flush=True)
    res["cases"].append(case)
    return case


def bench(M, N, K, iters=20, resid=False, act=ops.ACT_NONE):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    if resid:
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
        kw = dict(out_kind=ops.OUT_RESID_F32, out=out, gate=torch.randn(M // 768 + 1, N, device=dev), gate_rows=768,
                  out2=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw = dict(out=out, act=act)
    for _ in range(3):
        ops.gemm(a, w, b, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, b, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for comparison
    for _ in range(3):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    r = {"M": M, "N": N, "K": K, "resid": resid, "act": act, "ms": ms, "tflops": tf, "cublas_ms": ms2,
         "cublas_tflops": 2.0 * M * N * K / ms2 / 1e9}
    print(r, flush=True)
    return r


try:
    run_case(128, 128, 64, bias=False)
    run_case(128, 256, 64, bias=False)
    run_case(128, 256, 256)
    run_case(256, 512, 1024)
    run_case(1232, 1024, 768)                      # M tail (16 x 77 context rows)
    run_case(16, 6144, 1024, out_kind=ops.OUT_F32, act=ops.ACT_NONE)   # adaLN shape
    run_case(1536, 3072, 1024)
    run_case(1536, 4096, 1024, act=ops.ACT_GELU_ERF)
    run_case(1536, 1024, 4096, out_kind=ops.OUT_RESID_F32, gate=True, out2=True)
    run_case(1536, 1024, 1024, out_kind=ops.OUT_RESID_F32, gate=False)
    run_case(300, 1024, 256, act=ops.ACT_SILU)
    run_case(12288, 3072, 1024)
    res["bench"] = [bench(12288, 3072, 1024), bench(12288, 4096, 1024, act=ops.ACT_GELU_ERF),
                    bench(12288, 1024, 4096, resid=True), bench(12288, 1024, 1024, resid=True),
                    bench(12288, 1024, 1024)]
    res["ok"] = all(c["rel_l2"] < 1e-2 for c in res["cases"])
except Exception as e:  # noqa
    res["error"] = repr(e)
    res["ok"] = False
    print("ERROR", repr(e), flush=True)

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/gemm_check.json", "w") as f:
    json.dump(res, f, indent=1)
print("OK" if res.get("ok") else "FAILED")
