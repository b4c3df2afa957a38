"""TEST INFRASTRUCTURE ONLY (oracle) -- fp32 CPU restatement of the tri-plane VAE decode
(latent (B,12,32,32) -> tri-plane (B,96,128,128)).  Only tests/, smoke() and bench.py's CPU legs may
import this module.

Functional, on an AE-decoder state_dict with the reference's key names:
  vit/vit_triplane.py:58-108     PatchEmbedTriplane (grouped conv + channel interleave)   (V1)
  vit/vit_triplane.py:996-1011   vit_decode_backbone (ldm_upsample -> vit_decoder)
  dit/dit_decoder.py:15-151      DiT2 / DiTBlock2 / modulate2 (per-token adaLN, alternating
                                 in-plane / global attention)                               (V2)
  vit/vit_triplane.py:1913-1977  vit_decode_postprocess (unflatten -> conv_sr -> plane-major)
  ldm/modules/diffusionmodules/model.py:46-69,94-153,209-272,625-731  Decoder             (V3)
Pinned by oracle/make_golden.py against the reference's own modules (tests/golden/decoder.npz).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .dit import fused_mlp, layer_norm, self_attention

DIT2_SIZES = {"DiT2-S/2": dict(depth=12, hidden=384, heads=6), "DiT2-B/2": dict(depth=12, hidden=768, heads=12),
              "DiT2-L/2": dict(depth=24, hidden=1024, heads=16)}


def patch_embed_triplane(sd, latent: torch.Tensor, pre="superresolution.ldm_upsample.") -> torch.Tensor:
    B = latent.shape[0]
    y = F.conv2d(latent, sd[pre + "proj.weight"], sd[pre + "proj.bias"], stride=2, groups=3)
    y = y.reshape(B, y.shape[1] // 3, 3, y.shape[-2], y.shape[-1])  # B C 3 H W
    return y.flatten(2).transpose(1, 2)  # B (3 H W) C


def dit2_forward(sd, arch: str, c: torch.Tensor, pre="vit_decoder.") -> torch.Tensor:
    cfg = DIT2_SIZES[arch]
    heads = cfg["heads"]
    B = c.shape[0]
    x = sd[pre + "pos_embed"].repeat(B, 1, 1)
    sc = F.silu(c)
    for i in range(cfg["depth"]):
        p = f"{pre}blocks.{i}."
        mod = F.linear(sc, sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=-1)
        h = layer_norm(x) * (1 + sc_a) + sh_a
        if i % 2 == 0:  # attention within each plane: 'b (n l) c -> (b n) l c'
            T, D = h.shape[1], h.shape[2]
            a = self_attention(sd, p + "attn.", h.reshape(B * 3, T // 3, D), heads).reshape(B, T, D)
        else:
            a = self_attention(sd, p + "attn.", h, heads)
        x = x + g_a * a
        x = x + g_m * fused_mlp(sd, p + "mlp.", layer_norm(x) * (1 + sc_m) + sh_m)
    return x


def _gn(x, w, b):
    return F.group_norm(x, 32, w, b, eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _resblock(sd, p, x):
    h = F.conv2d(_swish(_gn(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])), sd[p + "conv1.weight"],
                 sd[p + "conv1.bias"], padding=1)
    h = F.conv2d(_swish(_gn(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"])), sd[p + "conv2.weight"],
                 sd[p + "conv2.bias"], padding=1)
    if (p + "nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def _attnblock(sd, p, x):
    h = _gn(x, sd[p + "norm.weight"], sd[p + "norm.bias"])
    q, k, v = (F.conv2d(h, sd[p + n + ".weight"], sd[p + n + ".bias"]) for n in ("q", "k", "v"))
    B, C, H, W = q.shape
    q, k, v = (t.reshape(B, C, H * W).transpose(1, 2) for t in (q, k, v))
    a = torch.softmax(q @ k.transpose(1, 2) * (C ** -0.5), dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, C, H, W)
    return x + F.conv2d(a, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def ldm_decoder(sd, z: torch.Tensor, pre="superresolution.conv_sr.", num_resolutions=4, num_res_blocks=1):
    h = F.conv2d(z, sd[pre + "conv_in.weight"], sd[pre + "conv_in.bias"], padding=1)
    h = _resblock(sd, pre + "mid.block_1.", h)
    h = _attnblock(sd, pre + "mid.attn_1.", h)
    h = _resblock(sd, pre + "mid.block_2.", h)
    for lvl in reversed(range(num_resolutions)):
        for b in range(num_res_blocks + 1):
            h = _resblock(sd, f"{pre}up.{lvl}.block.{b}.", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"{pre}up.{lvl}.upsample.conv.weight"], sd[f"{pre}up.{lvl}.upsample.conv.bias"], padding=1)
    h = _swish(_gn(h, sd[pre + "norm_out.weight"], sd[pre + "norm_out.bias"]))
    return F.conv2d(h, sd[pre + "conv_out.weight"], sd[pre + "conv_out.bias"], padding=1)


def vae_decode(sd: dict, arch: str, latent: torch.Tensor, scaling_divider: float = 1.0) -> torch.Tensor:
    """latent (B,12,32,32) -> tri-plane (B,96,128,128), channel = plane*32 + c."""
    sd = {k: v.float() for k, v in sd.items()}
    lat = latent.float() * scaling_divider                       # train_util_diffusion.py:188
    tok = patch_embed_triplane(sd, lat)                          # (B, 768, D)
    tok = dit2_forward(sd, arch, tok)
    B, L, C = tok.shape
    hw = int(round((L // 3) ** 0.5))
    z = tok.reshape(B, 3, hw, hw, C).permute(0, 1, 4, 2, 3).reshape(B * 3, C, hw, hw)   # '(b n) c h w'
    y = ldm_decoder(sd, z)                                       # (3B, 32, 128, 128)
    return y.reshape(B, 3 * y.shape[1], y.shape[2], y.shape[3])  # 'b (n c) h w'
