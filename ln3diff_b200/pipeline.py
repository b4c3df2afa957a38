"""Fused generation pipelines over the mirrored modules (the calls the reference's engines make).

`sample_t23d` is `DiffusionEngineLSGM.sample` (nsr/lsgm/sgm_DiffusionEngine.py:385-407):
EulerEDMSampler(num_steps) + DiscreteDenoiser(EpsScaling, 1000 idx) + VanillaCFG(scale) around the
T23D DiT.  Everything that is step-invariant is hoisted out of the loop -- the sigma schedule, the
nearest-of-1000 quantisation (two (1000, 2B) abs-diff argmins per step in the reference,
denoiser.py:64-78), c_in / c_out, the update coefficients -- into small device tables, so one step is
    DiT forward (2B samples, c_in folded into the patch embed)  +  1 fused update launch:
    x' = x + (dt/s) * sq * [(1-g) net_u + g net_c]        (sq = quantised sigma, c_out = -sq)
which is algebraically the reference's  D = x - sq*net;  x' = x + (x - cfg(D))/s * dt.
"""
from __future__ import annotations

import torch

from . import ops
from .sgm.modules.diffusionmodules.discretizer import LegacyDDPMDiscretization


@torch.no_grad()
def edm_cfg_tables(num_steps: int, scale: float, B: int, device, dtype=torch.float32):
    """Per-step device tables for the fused Euler-EDM + CFG loop (float32 arithmetic in the same
    order as the reference's scalar tensors)."""
    disc = LegacyDDPMDiscretization()
    sigmas = disc(num_steps, device="cpu")                       # (num_steps+1,), last = 0
    table = disc(1000, do_append_zero=False, flip=True)          # ascending, DiscreteDenoiser.sigmas
    s = sigmas[:-1]
    idx = (s[None, :] - table[:, None]).abs().argmin(dim=0)      # sigma_to_idx
    sq = table[idx]                                              # quantised sigma
    idx2 = (sq[None, :] - table[:, None]).abs().argmin(dim=0)    # possibly_quantize_c_noise
    c_in = 1 / (sq ** 2 + 1.0) ** 0.5
    r = (sigmas[1:] - s) / s                                     # dt / sigma
    w_u = r * sq * (1 - scale)
    w_c = r * sq * scale
    coef = torch.stack([torch.ones_like(r), w_u, w_c, torch.zeros_like(r)], 1)  # (steps, 4)
    return dict(
        init_scale=float(torch.sqrt(1.0 + sigmas[0] ** 2.0)),
        t_idx=idx2.to(device=device, dtype=torch.float32)[:, None].repeat(1, 2 * B).contiguous(),
        c_in=c_in.to(device)[:, None].repeat(1, 2 * B).contiguous(),
        coef=coef.to(device)[:, None, :].repeat(1, B, 1).contiguous(),
        sigmas=sigmas,
    )


@torch.no_grad()
def sample_t23d(model, randn: torch.Tensor, c: dict, uc: dict, num_steps: int = 250,
                scale: float = 6.5, tables: dict | None = None, use_graph: bool = True) -> torch.Tensor:
    """randn (B, 12, 32, 32) fp32 on the GPU (the reference draws it on the CPU generator and moves
    it, sgm_DiffusionEngine.py:395); c / uc = {'crossattn': (B, 77, ctx_dim)}.  Returns the
    denoised latents (B, 12, 32, 32) fp32."""
    if not randn.is_cuda:
        raise RuntimeError("sample_t23d runs on CUDA only (no CPU fallback)")
    B = randn.shape[0]
    if tables is None:
        tables = edm_cfg_tables(num_steps, scale, B, randn.device)
    ctx = torch.cat((uc["crossattn"], c["crossattn"]), 0).contiguous()   # VanillaCFG order: (uc, c)
    x = (randn.float() * tables["init_scale"]).contiguous()
    xa, xb = x, torch.empty_like(x)
    if use_graph and num_steps >= 4:
        # one CUDA graph = one DiT forward of the 2B CFG batch; replayed every step
        g = model.capture_graph(2 * B, ctx)
        for i in range(num_steps):
            g.x[:B].copy_(xa)
            g.x[B:].copy_(xa)
            g.t.copy_(tables["t_idx"][i])
            g.in_scale.copy_(tables["c_in"][i])
            g.replay()
            ops.sampler_affine_update(xa, tables["coef"][i], g.out[:B], g.out[B:], out=xb)
            xa, xb = xb, xa
        return xa.clone()
    x2 = torch.empty((2 * B,) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
    for i in range(num_steps):
        x2[:B].copy_(xa)
        x2[B:].copy_(xa)
        net = model(x2, tables["t_idx"][i], ctx, in_scale=tables["c_in"][i])
        ops.sampler_affine_update(xa, tables["coef"][i], net[:B], net[B:], out=xb)
        xa, xb = xb, xa
    return xa
