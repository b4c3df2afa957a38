"""GPU stand-in for "the reference GPU PyTorch/xformers path" of BASELINE.json's north star.

The reference tree cannot travel to the GPU box (and its GPU path needs xformers / timm, which are not
in the image), so this file restates -- in stock eager PyTorch, nothing of libln3b200 -- WHAT the
reference executes per sampling step on a GPU, with the reference's own structure and inefficiencies
left in (they are part of the baseline being compared against):

  * fp32 nn.Parameters under one `torch.autocast(bf16)` region (nsr/lsgm/sgm_DiffusionEngine.py:385-407),
    so every Linear runs a bf16 cuBLAS GEMM and LayerNorm runs in fp32;
  * `clip_text_proj(context)` re-evaluated in every forward (dit/dit_trilatent.py:107) and `to_k/to_v`
    of the constant context re-projected in every block of every step with three `.contiguous()` head
    permutes (ldm/modules/attention.py:278-296);
  * per-block `adaLN_modulation(c)` (dit/dit_models_xformers.py:285-294), separate modulate / gate /
    residual elementwise launches;
  * xformers `memory_efficient_attention` -> `F.scaled_dot_product_attention` (flash backend), xformers
    FusedMLP -> Linear + bias + exact GELU + Linear;
  * the sgm sampler as written: two (1000 x 2B) abs-diff argmins per step for the sigma quantisation
    (sgm/modules/diffusionmodules/denoiser.py:64-78), `torch.cat([x]*2)`, chunk, CFG and Euler update as
    separate launches (sampling.py:96-130, guiders.py:24-42).

`bench.py` times this in the same process, before the repo's arm (SURVEY.md section 8d "Timing protocol") and
reports it as `gpu_reference`.  It is a baseline, never part of the product path.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class _Mlp(nn.Module):
    def __init__(self, i, h, o, approximate="none"):
        super().__init__()
        self.fc1, self.fc2, self.approximate = nn.Linear(i, h), nn.Linear(h, o), approximate

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x), approximate=self.approximate))


class _SelfAttention(nn.Module):            # vit/vision_transformer.py:62-124
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.heads, C // self.heads).unbind(2)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return self.proj(o.transpose(1, 2).reshape(B, N, C))


class _CrossAttention(nn.Module):           # ldm/modules/attention.py:245-307
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.Linear(dim, dim)

    def forward(self, x, context):
        q, k, v = self.to_q(x), self.to_k(context), self.to_v(context)
        b, h = q.shape[0], self.heads
        q, k, v = (t.unsqueeze(3).reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3)
                   .reshape(b * h, t.shape[1], -1).contiguous() for t in (q, k, v))
        o = F.scaled_dot_product_attention(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0))[0]
        o = o.unsqueeze(0).reshape(b, h, o.shape[1], -1).permute(0, 2, 1, 3).reshape(b, o.shape[1], -1)
        return self.to_out(o)


class _Block(nn.Module):                    # dit/dit_models_xformers.py:231-323 (TextCondDiTBlock)
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.attn = _SelfAttention(dim, heads)
        self.cross_attn = _CrossAttention(dim, dim, heads)
        self.fc1, self.fc2 = nn.Linear(dim, 4 * dim), nn.Linear(4 * dim, dim)
        self.ada = nn.Linear(dim, 6 * dim)

    def forward(self, x, c, context):
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = self.ada(F.silu(c)).chunk(6, dim=1)
        x = x + g_a.unsqueeze(1) * self.attn(self.norm1(x) * (1 + sc_a.unsqueeze(1)) + sh_a.unsqueeze(1))
        x = x + self.cross_attn(x, context)
        h = self.norm2(x) * (1 + sc_m.unsqueeze(1)) + sh_m.unsqueeze(1)
        return x + g_m.unsqueeze(1) * self.fc2(F.gelu(self.fc1(h)))


class EagerDiT(nn.Module):
    """DiT_TriLatent (T23D, roll_out) in stock PyTorch modules; weights copied from the mirror's state_dict."""

    def __init__(self, depth=24, dim=1024, heads=16, ctx_dim=768, in_ch=4):
        super().__init__()
        self.in_ch, self.dim = in_ch, dim
        self.x_embedder = nn.Conv2d(in_ch, dim, 2, 2)
        self.t0, self.t2 = nn.Linear(256, dim), nn.Linear(dim, dim)
        self.clip_text_proj = _Mlp(ctx_dim, dim, dim, approximate="tanh")
        self.pos_embed = nn.Parameter(torch.zeros(1, 768, dim), requires_grad=False)
        self.blocks = nn.ModuleList([_Block(dim, heads, ctx_dim) for _ in range(depth)])
        self.final_norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.final_ada, self.final_linear = nn.Linear(dim, 2 * dim), nn.Linear(dim, 4 * in_ch)

    @torch.no_grad()
    def load_mirror_state_dict(self, sd):
        cp = lambda dst, key: dst.copy_(sd[key].to(dst.device, dst.dtype))
        cp(self.x_embedder.weight, "x_embedder.proj.weight"); cp(self.x_embedder.bias, "x_embedder.proj.bias")
        cp(self.t0.weight, "t_embedder.mlp.0.weight"); cp(self.t0.bias, "t_embedder.mlp.0.bias")
        cp(self.t2.weight, "t_embedder.mlp.2.weight"); cp(self.t2.bias, "t_embedder.mlp.2.bias")
        for n in ("fc1", "fc2"):
            cp(getattr(self.clip_text_proj, n).weight, f"clip_text_proj.y_proj.{n}.weight")
            cp(getattr(self.clip_text_proj, n).bias, f"clip_text_proj.y_proj.{n}.bias")
        cp(self.pos_embed, "pos_embed")
        for i, b in enumerate(self.blocks):
            p = f"blocks.{i}."
            cp(b.attn.qkv.weight, p + "attn.qkv.weight"); cp(b.attn.qkv.bias, p + "attn.qkv.bias")
            cp(b.attn.proj.weight, p + "attn.proj.weight"); cp(b.attn.proj.bias, p + "attn.proj.bias")
            for n in ("to_q", "to_k", "to_v"):
                cp(getattr(b.cross_attn, n).weight, p + f"cross_attn.{n}.weight")
            cp(b.cross_attn.to_out.weight, p + "cross_attn.to_out.0.weight")
            cp(b.cross_attn.to_out.bias, p + "cross_attn.to_out.0.bias")
            cp(b.fc1.weight, p + "mlp.mlp.0.weight"); cp(b.fc1.bias, p + "mlp.mlp.1.bias")
            cp(b.fc2.weight, p + "mlp.mlp.2.weight"); cp(b.fc2.bias, p + "mlp.mlp.3.bias")
            cp(b.ada.weight, p + "adaLN_modulation.1.weight"); cp(b.ada.bias, p + "adaLN_modulation.1.bias")
        cp(self.final_ada.weight, "final_layer.adaLN_modulation.1.weight")
        cp(self.final_ada.bias, "final_layer.adaLN_modulation.1.bias")
        cp(self.final_linear.weight, "final_layer.linear.weight"); cp(self.final_linear.bias, "final_layer.linear.bias")
        return self

    def forward(self, x, timesteps, context):
        B, C3, H, W = x.shape
        c = C3 // 3
        half = 128
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=x.device) / half)
        args = timesteps[:, None].float() * freqs[None]
        t = self.t2(F.silu(self.t0(torch.cat([torch.cos(args), torch.sin(args)], dim=-1))))
        context = self.clip_text_proj(context)                                    # every step (dit_trilatent.py:107)
        xr = x.reshape(B, c, 3, H, W).permute(0, 2, 1, 3, 4).reshape(B * 3, c, H, W)
        h = self.x_embedder(xr).flatten(2).transpose(1, 2).reshape(B, 768, self.dim) + self.pos_embed
        for blk in self.blocks:
            h = blk(h, t, context)
        shift, scale = self.final_ada(F.silu(t)).chunk(2, dim=1)
        h = self.final_linear(self.final_norm(h) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1))
        out_ch = h.shape[-1] // 4
        h = h.reshape(B * 3, 16, 16, 2, 2, out_ch)
        h = torch.einsum("nhwpqc->nchpwq", h).reshape(B * 3, out_ch, 32, 32)
        return h.reshape(B, 3, out_ch, 32, 32).permute(0, 2, 1, 3, 4).reshape(B, out_ch * 3, 32, 32).float().contiguous()


def _legacy_ddpm_sigmas(n, device, append_zero=True, flip=False):
    betas = torch.linspace(0.00085 ** 0.5, 0.0120 ** 0.5, 1000, dtype=torch.float64).numpy() ** 2
    ac = np.cumprod(1.0 - betas, axis=0)
    if n < 1000:
        ac = ac[np.linspace(999, 0, n, endpoint=False).astype(int)[::-1]]
    sig = torch.flip(torch.tensor((1 - ac) / ac, dtype=torch.float32, device=device) ** 0.5, (0,))
    if append_zero:
        sig = torch.cat([sig, sig.new_zeros([1])])
    return torch.flip(sig, (0,)) if flip else sig


@torch.no_grad()
def euler_edm_cfg_steps(model, x, c, uc, num_steps, scale, first_steps=None):
    """The shipped T23D sampler as the reference's classes execute it (EulerEDMSampler + DiscreteDenoiser
    (EpsScaling) + VanillaCFG), eager, for the first `first_steps` of `num_steps` steps (all if None)."""
    dev = x.device
    table = _legacy_ddpm_sigmas(1000, dev, append_zero=False, flip=True)
    sigmas = _legacy_ddpm_sigmas(num_steps, dev)
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    c_cat = torch.cat((uc, c), 0)
    with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=x.is_cuda):
        for i in range(first_steps if first_steps is not None else num_steps):
            sigma, nxt = s_in * sigmas[i], s_in * sigmas[i + 1]
            xin, sin = torch.cat([x] * 2), torch.cat([sigma] * 2)
            sq = table[(sin - table[:, None]).abs().argmin(dim=0)]
            sq4 = sq[:, None, None, None]
            c_in = 1 / (sq4 ** 2 + 1.0) ** 0.5
            c_noise = (sq - table[:, None]).abs().argmin(dim=0)
            den = model(xin * c_in, c_noise, c_cat) * (-sq4) + xin * torch.ones_like(sq4)
            x_u, x_c = den.chunk(2)
            d = (x - (x_u + scale * (x_c - x_u))) / sigma[:, None, None, None]
            x = x + (nxt - sigma)[:, None, None, None] * d
    return x
