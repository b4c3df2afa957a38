"""GPU diagnostic: attention / elementwise kernels vs torch fp32, DiT forward vs the CPU oracle,
and a DiT-L/2 forward timing.  Writes gpurun_out/dit_check.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from ln3diff_b200 import ops
from ln3diff_b200._lib import NORM_LAYER, NORM_NONE, NORM_RMS
from ln3diff_b200.dit.dit_trilatent import DiT_models
from ln3diff_b200.dit.dit_models_xformers import TextCondDiTBlock
from oracle import dit as odit

dev = "cuda"
torch.manual_seed(0)
res = {"cases": {}}


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def record(name, val, tol):
    ok = bool(val < tol)
    res["cases"][name] = {"rel_l2": val, "tol": tol, "ok": ok}
    print(name, val, "OK" if ok else "FAIL", flush=True)


def stage(fn, name):
    try:
        fn()
    except Exception as e:  # noqa
        res["cases"][name] = {"error": repr(e), "ok": False}
        print(name, "ERROR", repr(e), flush=True)


def check_fmha():
    for (B, H, Lq, Lkv, tag) in [(2, 12, 768, 768, "self768"), (2, 16, 768, 77, "cross77"),
                                 (1, 4, 200, 333, "tails"), (2, 16, 768, 1024, "i23d_kv1024"),
                                 (3, 16, 256, 256, "plane256")]:
        D = H * 64
        qkv = (torch.randn(B, max(Lq, Lkv), 3 * D, device=dev)).bfloat16()
        q = qkv[:, :Lq, 0:D]
        k = qkv[:, :Lkv, D:2 * D]
        v = qkv[:, :Lkv, 2 * D:]
        out = ops.fmha(q, k, v, H)
        torch.cuda.synchronize()
        qf, kf, vf = (t.float().reshape(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
        ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, Lq, D)
        record("fmha_" + tag, rel(out, ref), 1e-2)


def check_elementwise():
    x = torch.randn(300, 1024, device=dev) * 2 + 0.5
    sh = torch.randn(3, 6 * 1024, device=dev)
    out = ops.norm_modulate(x, norm=NORM_LAYER, shift=sh[:, 0:1024], scale=sh[:, 1024:2048], mod_rows=100)
    ref = F.layer_norm(x, (1024,), eps=1e-6) * (1 + sh[:, 1024:2048].repeat_interleave(100, 0)) + \
        sh[:, 0:1024].repeat_interleave(100, 0)
    record("ln_modulate", rel(out, ref), 4e-3)
    w = torch.randn(768, device=dev)
    x7 = torch.randn(77, 768, device=dev)
    out = ops.norm_modulate(x7, norm=NORM_RMS, weight=w, eps=1e-5)
    ref = x7 * torch.rsqrt(x7.pow(2).mean(-1, keepdim=True) + 1e-5) * w
    record("rms_norm", rel(out, ref), 4e-3)
    out = ops.norm_modulate(x7, norm=NORM_NONE, act=ops.ACT_SILU)
    record("cast_silu", rel(out, F.silu(x7)), 4e-3)
    t = torch.tensor([0.0, 1.0, 17.0, 999.0, 0.37], device=dev)
    out = ops.timestep_embedding(t)
    record("timestep_embedding", rel(out, odit.timestep_embedding(t.cpu()).to(dev)), 4e-3)
    # patch embed / final layer vs oracle helpers
    sd = {"x_embedder.proj.weight": torch.randn(768, 4, 2, 2), "x_embedder.proj.bias": torch.randn(768)}
    xin = torch.randn(2, 12, 32, 32)
    pos = torch.randn(1, 768, 768)
    ref = odit.patch_embed_rollout(sd, xin) + pos
    out = ops.patch_embed(xin.to(dev), sd["x_embedder.proj.weight"].to(dev),
                          sd["x_embedder.proj.bias"].to(dev), pos.to(dev))
    record("patch_embed", rel(out.cpu(), ref), 1e-5)
    tok = torch.randn(2, 768, 768)
    shift, scale = torch.randn(2, 768), torch.randn(2, 768)
    wf, bfin = torch.randn(16, 768) * 0.05, torch.randn(16)
    h = odit.layer_norm(tok) * (1 + scale[:, None]) + shift[:, None]
    ref = odit.unpatchify_rollout(F.linear(h, wf, bfin), 4)
    out = ops.final_layer(tok.to(dev), shift.to(dev), scale.to(dev), wf.to(dev), bfin.to(dev), 32)
    record("final_layer", rel(out.cpu(), ref), 1e-5)
    x = torch.randn(3, 12, 32, 32, device=dev)
    m0, m1, nz = (torch.randn_like(x) for _ in range(3))
    cf = torch.randn(3, 4, device=dev)
    out = ops.sampler_affine_update(x, cf, m0, m1, nz)
    c = cf[:, :, None, None, None]
    ref = c[:, 0] * x + c[:, 1] * m0 + c[:, 2] * m1 + c[:, 3] * nz
    record("sampler_update", rel(out, ref), 1e-6)


def build(arch):
    torch.manual_seed(0)
    m = DiT_models[arch](input_size=32, num_classes=0, learn_sigma=False, in_channels=4,
                         context_dim=768, roll_out=True, vit_blk=TextCondDiTBlock)
    sd = odit.derandomize_zero_init(m.state_dict())
    m.load_state_dict(sd)
    return m, sd


def check_dit():
    m, sd = build("DiT-B/2")
    x = torch.randn(2, 12, 32, 32)
    t = torch.tensor([10, 500])
    ctx = torch.randn(2, 77, 768)
    t0 = time.time()
    ref = odit.dit_t23d_forward(sd, "DiT-B/2", x, t, ctx)
    res["oracle_ditb_fwd_s"] = time.time() - t0
    m = m.to(dev)
    out = m(x.to(dev), t.to(dev), {"crossattn": ctx.to(dev)})
    torch.cuda.synchronize()
    record("dit_b2_forward_vs_oracle", rel(out.cpu(), ref), 2e-2)
    res["dit_b2_absmax"] = [out.abs().max().item(), ref.abs().max().item()]


def time_dit():
    m, _ = build("DiT-L/2")
    m = m.to(dev)
    B = 16
    x = torch.randn(B, 12, 32, 32, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)
    ctx = torch.randn(B, 77, 768, device=dev)
    for _ in range(3):
        m(x, t, ctx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 10
    w0 = time.time()
    e0.record()
    for _ in range(n):
        m(x, t, ctx)
    e1.record()
    host_issue = (time.time() - w0) / n
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res["dit_l2_b16_forward_ms"] = ms
    res["dit_l2_b16_host_issue_ms"] = host_issue * 1e3
    res["dit_l2_b16_tflops"] = 0.613 * B / ms * 1e3
    print("DiT-L/2 B'=16 forward", ms, "ms;", res["dit_l2_b16_tflops"], "TF/s; host issue",
          host_issue * 1e3, "ms", flush=True)


stage(check_elementwise, "elementwise")
stage(check_fmha, "fmha")
stage(check_dit, "dit")
stage(time_dit, "time_dit")
res["ok"] = all(c.get("ok", False) for c in res["cases"].values())
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/dit_check.json", "w") as f:
    json.dump(res, f, indent=1)
print("OK" if res["ok"] else "FAILED")
