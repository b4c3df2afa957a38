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

import os

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
        # every sample of a step shares its timestep: the adaLN modulations of all steps in one pass
        shared = hasattr(model, "modulation_table") and os.environ.get("LN3_SHARED_MODULATION", "1") != "0"
        g = model.capture_graph(2 * B, ctx, shared_mod=True) if shared else model.capture_graph(2 * B, ctx)
        mod_table = model.modulation_table(tables["t_idx"][:num_steps, 0]) if shared else None
        for i in range(num_steps):
            g.x[:B].copy_(xa)
            g.x[B:].copy_(xa)
            if shared:
                g.mod.copy_(mod_table[i:i + 1])
            else:
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


@torch.no_grad()
def decode_and_render(decoder, latents: torch.Tensor, cameras: torch.Tensor, resolution: int = 128,
                      scaling_divider: float = 0.96806, noise: tuple | None = None, mlp_tf32: bool = True):
    """`TrainLoopDiffusionWithRec.render_video_given_triplane` (nsr/train_util_diffusion.py:176-382)
    without the host round trips: latents (B,12,32,32) -> tri-planes (decoded ONCE; the reference
    decodes twice, :204-206 and :268-270) -> every camera of `cameras` (V,25) for every latent in
    one fused renderer launch.  Per-view global reductions (group_size=1) reproduce the reference's
    one-view-per-call loop (:292-302).  Returns image_raw (B,V,3,H,W) in [-1,1], image_depth
    (B,V,1,H,W), image_mask (B,V,1,H,W).  `noise` = (coarse, fine) tensors of shape (B*V, H*W, 64) to
    override the device RNG (tests)."""
    if not latents.is_cuda:
        raise RuntimeError("decode_and_render runs on CUDA only (no CPU fallback)")
    B, V = latents.shape[0], cameras.shape[0]
    planes_cl = decoder.decode_to_channels_last(latents, in_mul=scaling_divider)     # (B,3,128,128,32)
    cams = cameras.to(latents.device, torch.float32).repeat(B, 1).contiguous()       # (B*V, 25)
    ray_o, ray_d = ops.generate_rays(cams, resolution)
    M = resolution * resolution
    if noise is None:
        noise = (torch.rand(B * V, M, 64, device=latents.device), torch.rand(B * V, M, 64, device=latents.device))
    kw = decoder.rendering_kwargs
    out = ops.render_views(planes_cl, ray_o, ray_d, noise[0].contiguous(), noise[1].contiguous(),
                           decoder.triplane_decoder.decoder.raw_parameters(), views_per_obj=V, group_size=1,
                           box_warp=kw.get("box_warp", 0.9), bbox_min=kw.get("sampler_bbox_min", -0.45),
                           bbox_max=kw.get("sampler_bbox_max", 0.45), white_back=kw.get("white_back", True),
                           mlp_tf32=mlp_tf32)
    H = W = resolution
    w = out["weights"].view(B, V, 1, H, W)
    return dict(image_raw=out["rgb"].view(B, V, 3, H, W), image_depth=out["depth"].view(B, V, 1, H, W),
                weights_samples=w, image_mask=w * (1 + 2 * 0.001) - 0.001)


@torch.no_grad()
def generate_t23d(model, decoder, randn, c, uc, cameras, num_steps: int = 250, scale: float = 6.5,
                  resolution: int = 128):
    """Text-to-3D end to end on one GPU: sample -> decode -> render (the body of
    DiffusionEngineLSGM.eval_cldm, nsr/lsgm/sgm_DiffusionEngine.py:410-523, minus conditioner + video sink)."""
    latents = sample_t23d(model, randn, c, uc, num_steps, scale)
    return latents, decode_and_render(decoder, latents, cameras, resolution)
