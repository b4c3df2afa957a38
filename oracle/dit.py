"""TEST INFRASTRUCTURE ONLY (oracle) -- fp32 CPU restatement of the reference DiT denoisers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (ln3diff_b200) never does.

Functional restatement (weights come in as a reference-named state_dict) of
  dit/dit_trilatent.py:74-143        DiT_TriLatent.forward (T23D, roll_out)
  dit/dit_models_xformers.py:87-127  TimestepEmbedder
  dit/dit_models_xformers.py:183-223 CaptionEmbedder (timm Mlp, tanh-GELU)
  dit/dit_models_xformers.py:231-323 DiTBlock / TextCondDiTBlock (adaLN-Zero)
  vit/vision_transformer.py:106-124  MemEffAttention (qkv -> softmax(qk^T/sqrt d) v -> proj)
  ldm/modules/attention.py:245-307   MemoryEfficientCrossAttention (no bias on q/k/v, no pre-norm)
  dit/dit_models_xformers.py:655-678 FinalLayer, :821-835 unpatchify
Pinned against the reference's own modules (imported with oracle/_stubs.py) by
oracle/make_golden.py -> tests/golden/dit_t23d_*.npz; third-party arithmetic (xformers FMHA /
FusedMLP, timm PatchEmbed/Mlp) is "parity unpinned" beyond that.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

DIT_SIZES = {  # dit/dit_models_xformers.py:1029-1106 ; dit/dit_trilatent.py:270-327
    "DiT-B/2": dict(depth=12, hidden=768, heads=12),
    "DiT-L/2": dict(depth=24, hidden=1024, heads=16),
    "DiT-XL/2": dict(depth=28, hidden=1152, heads=16),
    "DiT-S/2": dict(depth=12, hidden=384, heads=6),
}


def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0) -> torch.Tensor:
    """dit_models_xformers.py:97-121: [cos(t f), sin(t f)], f_i = exp(-ln(1e4) i / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def layer_norm(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps)


def rms_norm(x: torch.Tensor, weight: torch.Tensor | None, eps: float = 1e-5) -> torch.Tensor:
    """dit/norm.py:27-40."""
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return y * weight if weight is not None else y


def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """q (B,H,Lq,d), k/v (B,H,Lk,d): softmax(q k^T / sqrt d) v  (xformers FMHA semantics)."""
    s = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    return s.softmax(-1) @ v


def self_attention(sd, pre: str, x: torch.Tensor, heads: int, qk_norm: bool = False) -> torch.Tensor:
    B, N, C = x.shape
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B, N, 3, heads, C // heads)
    q, k, v = qkv.unbind(2)  # (B,N,H,d)
    if qk_norm:
        q = rms_norm(q, sd[pre + "q_norm.weight"])
        k = rms_norm(k, sd[pre + "k_norm.weight"])
    o = sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def cross_attention(sd, pre: str, x: torch.Tensor, ctx: torch.Tensor, heads: int,
                    qk_norm: bool = False) -> torch.Tensor:
    B, N, _ = x.shape
    q = F.linear(x, sd[pre + "to_q.weight"])
    k = F.linear(ctx, sd[pre + "to_k.weight"])
    v = F.linear(ctx, sd[pre + "to_v.weight"])
    d = q.shape[-1] // heads
    q, k, v = (t.reshape(B, t.shape[1], heads, d).transpose(1, 2) for t in (q, k, v))
    if qk_norm:
        q = rms_norm(q, sd[pre + "q_norm.weight"])
        k = rms_norm(k, sd[pre + "k_norm.weight"])
    o = sdpa(q, k, v).transpose(1, 2).reshape(B, N, heads * d)
    return F.linear(o, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])


def fused_mlp(sd, pre: str, x: torch.Tensor) -> torch.Tensor:
    """xformers FusedMLP at eval: Linear(no bias) + bias -> erf-GELU -> Linear(no bias) + bias."""
    h = F.gelu(F.linear(x, sd[pre + "mlp.0.weight"]) + sd[pre + "mlp.1.bias"])
    return F.linear(h, sd[pre + "mlp.2.weight"]) + sd[pre + "mlp.3.bias"]


def patch_embed_rollout(sd, x: torch.Tensor, patch: int = 2) -> torch.Tensor:
    """'b (c n) h w -> (b n) c h w' -> Conv2d(k=s=patch) -> '(b n) l c -> b (n l) c'
    (dit_trilatent.py:93-99)."""
    B, C3, H, W = x.shape
    c = C3 // 3
    xr = x.reshape(B, c, 3, H, W).permute(0, 2, 1, 3, 4).reshape(B * 3, c, H, W)
    y = F.conv2d(xr, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=patch)
    D = y.shape[1]
    y = y.flatten(2).transpose(1, 2)  # (B*3, L, D)
    return y.reshape(B, 3 * y.shape[1], D)


def unpatchify_rollout(x: torch.Tensor, out_ch: int, patch: int = 2) -> torch.Tensor:
    """'b (n l) c -> (b n) l c' -> unpatchify -> '(b n) c h w -> b (c n) h w'
    (dit_trilatent.py:130-140, dit_models_xformers.py:821-835)."""
    B, T, _ = x.shape
    L = T // 3
    h = w = int(round(L ** 0.5))
    x = x.reshape(B * 3, h, w, patch, patch, out_ch)
    x = torch.einsum("nhwpqc->nchpwq", x).reshape(B * 3, out_ch, h * patch, w * patch)
    return x.reshape(B, 3, out_ch, h * patch, w * patch).permute(0, 2, 1, 3, 4).reshape(
        B, out_ch * 3, h * patch, w * patch)


def dit_t23d_forward(sd: dict, arch: str, x: torch.Tensor, timesteps: torch.Tensor,
                     context: torch.Tensor, first_blocks: int | None = None) -> torch.Tensor:
    """DiT_TriLatent(vit_blk=TextCondDiTBlock, FinalLayer).forward, fp32.

    x (B,12,32,32); timesteps (B,) int64 index / float; context (B,77,ctx_dim) -> (B,12,32,32).
    `first_blocks` runs only the first n transformer blocks (bench.py's bounded CPU-timing sample; every
    block costs the same) -- never used by a parity test.
    """
    cfg = DIT_SIZES[arch]
    heads, depth = cfg["heads"], cfg["depth"]
    if first_blocks is not None:
        depth = min(depth, first_blocks)
    sd = {k: v.float() for k, v in sd.items()}
    x = x.float()
    t = timestep_embedding(timesteps)
    t = F.linear(F.silu(F.linear(t, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                 sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    h = patch_embed_rollout(sd, x) + sd["pos_embed"]
    ctx = F.linear(F.gelu(F.linear(context.float(), sd["clip_text_proj.y_proj.fc1.weight"],
                                   sd["clip_text_proj.y_proj.fc1.bias"]), approximate="tanh"),
                   sd["clip_text_proj.y_proj.fc2.weight"], sd["clip_text_proj.y_proj.fc2.bias"])
    st = F.silu(t)
    for i in range(depth):
        p = f"blocks.{i}."
        mod = F.linear(st, sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)
        h = h + g_a[:, None] * self_attention(
            sd, p + "attn.", layer_norm(h) * (1 + sc_a[:, None]) + sh_a[:, None], heads)
        h = h + cross_attention(sd, p + "cross_attn.", h, ctx, heads)
        h = h + g_m[:, None] * fused_mlp(
            sd, p + "mlp.", layer_norm(h) * (1 + sc_m[:, None]) + sh_m[:, None])
    mod = F.linear(st, sd["final_layer.adaLN_modulation.1.weight"],
                   sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=1)
    h = layer_norm(h) * (1 + scale[:, None]) + shift[:, None]
    h = F.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    out_ch = sd["final_layer.linear.weight"].shape[0] // 4
    return unpatchify_rollout(h, out_ch).contiguous()


def derandomize_zero_init(sd: dict, std: float = 0.02, seed: int = 1234) -> dict:
    """Every all-zero floating tensor <- N(0, std^2) (adaLN-Zero / zero final layer would make
    every parity test vacuous: dit_models_xformers.py:807-819)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        if v.is_floating_point() and v.numel() > 0 and float(v.abs().max()) == 0.0:
            v = torch.randn(v.shape, generator=g, dtype=torch.float32) * std
        out[k] = v
    return out


def synth_state_dict(shapes: dict, seed: int = 7, keep: dict | None = None) -> dict:
    """Deterministic weights independent of module construction order: every tensor is drawn from
    a CPU generator in sorted-key order (weights N(0, 1/fan_in), biases / vectors N(0, 0.02^2));
    `keep` entries (e.g. the analytic pos_embed) are passed through."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes.keys()):
        if keep and k in keep:
            sd[k] = keep[k].clone()
            continue
        shp = tuple(shapes[k])
        if len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            sd[k] = torch.randn(shp, generator=g) * (1.0 / fan_in ** 0.5)
        else:
            sd[k] = torch.randn(shp, generator=g) * 0.02
    return sd


I23D_SIZES = {"DiT-PixArt-L/2": dict(depth=24, hidden=1024, heads=16),
              "DiT-PixArt-B/2": dict(depth=12, hidden=768, heads=12)}


def dit_i23d_pixart_forward(sd: dict, arch: str, x: torch.Tensor, timesteps: torch.Tensor,
                            context: dict) -> torch.Tensor:
    """DiT_I23D_PixelArt(vit_blk=ImageCondDiTBlockPixelArtRMSNorm, T2IFinalLayer).forward, fp32.

    Restates dit/dit_i23d.py:173-290 (model), dit/dit_models_xformers.py:481-539,604-618 (block with
    the shared adaLN + per-block scale_shift_table, RMSNorm pre-norms), :61-84 (T2IFinalLayer),
    vit/vision_transformer.py:106-124 with qk_norm, ldm/modules/attention.py:278-307 with qk_norm.
    context = {'vector': (B, pooling_ctx_dim), 'crossattn': (B, 256, 1024 clip | 1024 dino)}."""
    cfg = I23D_SIZES[arch]
    heads, depth = cfg["heads"], cfg["depth"]
    sd = {k: v.float() for k, v in sd.items()}
    x = x.float()
    B = x.shape[0]
    vec, ca = context["vector"].float(), context["crossattn"].float()
    cls = F.linear(F.layer_norm(vec, (vec.shape[-1],), sd["cap_embedder.0.weight"], sd["cap_embedder.0.bias"], 1e-5),
                   sd["cap_embedder.1.weight"], sd["cap_embedder.1.bias"])
    clip = rms_norm(ca[..., :1024], sd["attention_y_norm.weight"], 1e-5)
    dino = F.linear(F.gelu(F.linear(ca[..., 1024:], sd["dino_proj.y_proj.fc1.weight"], sd["dino_proj.y_proj.fc1.bias"]),
                           approximate="tanh"), sd["dino_proj.y_proj.fc2.weight"], sd["dino_proj.y_proj.fc2.bias"])
    t = timestep_embedding(timesteps.float())
    t = F.linear(F.silu(F.linear(t, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                 sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"]) + cls
    t0 = F.linear(F.silu(t), sd["adaLN_modulation.1.weight"], sd["adaLN_modulation.1.bias"])
    h = patch_embed_rollout(sd, x) + sd["pos_embed"]
    T = h.shape[1]
    for i in range(depth):
        p = f"blocks.{i}."
        mod = sd[p + "scale_shift_table"][None] + t0.reshape(B, 6, -1)
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)  # each (B,1,D)
        a_in = torch.cat([rms_norm(h, sd[p + "norm1.weight"], 1e-5) * (1 + sc_a) + sh_a, dino], dim=1)
        h = h + g_a * self_attention(sd, p + "attn.", a_in, heads, qk_norm=True)[:, :T]
        h = h + cross_attention(sd, p + "cross_attn.", h, clip, heads, qk_norm=True)
        h = h + g_m * fused_mlp(sd, p + "mlp.", rms_norm(h, sd[p + "norm2.weight"], 1e-5) * (1 + sc_m) + sh_m)
    shift, scale = (sd["final_layer.scale_shift_table"][None] + t[:, None]).chunk(2, dim=1)
    h = layer_norm(h) * (1 + scale) + shift
    h = F.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    return unpatchify_rollout(h, sd["final_layer.linear.weight"].shape[0] // 4).contiguous()


def dit_t23d_pixart_forward(sd: dict, arch: str, x: torch.Tensor, timesteps: torch.Tensor,
                            context: dict) -> torch.Tensor:
    """DiT_TriLatent_PixelArt (dit/dit_trilatent.py:146-246) with PixelArtTextCondDiTBlock
    (dit/dit_models_xformers.py:326-369) and T2IFinalLayer (:61-84), fp32.
    arch in {'DiT-PixelArt-L/2', 'DiT-PixelArt-B/2'}; context = {'vector': (B, Cc), 'crossattn': (B, L, Cc)}."""
    cfg = DIT_SIZES[arch.replace("-PixelArt", "")]
    heads, depth = cfg["heads"], cfg["depth"]
    sd = {k: v.float() for k, v in sd.items()}
    x = x.float()
    B = x.shape[0]
    vec, ca = context["vector"].float(), context["crossattn"].float()
    cls = F.linear(F.layer_norm(vec, (vec.shape[-1],), sd["cap_embedder.0.weight"], sd["cap_embedder.0.bias"], 1e-5),
                   sd["cap_embedder.1.weight"], sd["cap_embedder.1.bias"])
    t = timestep_embedding(timesteps.float())
    t = F.linear(F.silu(F.linear(t, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                 sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"]) + cls
    t0 = F.linear(F.silu(t), sd["adaLN_modulation.1.weight"], sd["adaLN_modulation.1.bias"])
    h = patch_embed_rollout(sd, x) + sd["pos_embed"]
    for i in range(depth):
        p = f"blocks.{i}."
        mod = sd[p + "scale_shift_table"][None] + t0.reshape(B, 6, -1)
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)
        h = h + g_a * self_attention(sd, p + "attn.", rms_norm(h, sd[p + "norm1.weight"], 1e-5) * (1 + sc_a) + sh_a, heads)
        h = h + cross_attention(sd, p + "cross_attn.", h, rms_norm(ca, sd[p + "attention_y_norm.weight"], 1e-5), heads)
        h = h + g_m * fused_mlp(sd, p + "mlp.", rms_norm(h, sd[p + "norm2.weight"], 1e-5) * (1 + sc_m) + sh_m)
    shift, scale = (sd["final_layer.scale_shift_table"][None] + t[:, None]).chunk(2, dim=1)
    h = layer_norm(h) * (1 + scale) + shift
    h = F.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    return unpatchify_rollout(h, sd["final_layer.linear.weight"].shape[0] // 4).contiguous()

