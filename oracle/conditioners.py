"""TEST INFRASTRUCTURE ONLY (oracle) -- the conditioner towers in fp32 on the CPU.

Only tests/ and __graft_entry__.smoke() may import this module; the product path never does.

`FrozenCLIPEmbedder` (reference sgm/modules/encoders/modules.py:347-408) IS a `transformers.CLIPTextModel`, and
transformers is installed in this image: for the text tower the oracle is the reference's own dependency, run in
fp32 on random-init weights -- pinned.  `open_clip` and the dinov2 hub model are NOT in the image; for them the
oracle is transformers' port of the same published architectures (`CLIPVisionModelWithProjection`,
`Dinov2WithRegistersModel`), whose state dicts are renamed below to the key layout the reference objects hold
(`model.visual.*`, dinov2 hub names).  PARITY UNPINNED against open_clip / the hub code themselves; anchored on the
transformers ports.
"""
from __future__ import annotations

import torch


def _randomize_norms(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.LayerNorm):
                m.weight.copy_(1 + 0.2 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        for n, p in model.named_parameters():
            if n.endswith("bias") and p.abs().sum() == 0:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            if "lambda1" in n:
                p.copy_(0.5 + 0.2 * torch.randn(p.shape, generator=g))
    return model


def clip_text(depth=12, width=768, mlp=3072, vocab=49408, max_len=77, seed=0, init_scale=3.0):
    from transformers import CLIPTextConfig, CLIPTextModel
    torch.manual_seed(seed)
    cfg = CLIPTextConfig(vocab_size=vocab, hidden_size=width, intermediate_size=mlp, num_hidden_layers=depth,
                         num_attention_heads=width // 64, max_position_embeddings=max_len, hidden_act="quick_gelu",
                         eos_token_id=2, bos_token_id=0, pad_token_id=1, initializer_range=0.02 * init_scale)
    m = _randomize_norms(CLIPTextModel(cfg).eval().float(), seed)
    sd = {"transformer." + k: v for k, v in m.state_dict().items() if "position_ids" not in k}
    return m, sd


def clip_vision(depth=2, width=256, mlp=1024, embed=128, seed=0, init_scale=3.0):
    """-> (HF model, open_clip-style state dict with the `model.visual.` prefix)."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(seed)
    cfg = CLIPVisionConfig(hidden_size=width, intermediate_size=mlp, num_hidden_layers=depth, num_attention_heads=width // 64,
                           image_size=224, patch_size=14, hidden_act="quick_gelu", projection_dim=embed,
                           initializer_range=0.02 * init_scale)
    m = _randomize_norms(CLIPVisionModelWithProjection(cfg).eval().float(), seed)
    h = m.state_dict()
    v = "vision_model."
    sd = {"conv1.weight": h[v + "embeddings.patch_embedding.weight"], "class_embedding": h[v + "embeddings.class_embedding"],
          "positional_embedding": h[v + "embeddings.position_embedding.weight"],
          "ln_pre.weight": h[v + "pre_layrnorm.weight"], "ln_pre.bias": h[v + "pre_layrnorm.bias"],
          "ln_post.weight": h[v + "post_layernorm.weight"], "ln_post.bias": h[v + "post_layernorm.bias"],
          "proj": h["visual_projection.weight"].t().contiguous()}
    for i in range(depth):
        s, d = v + f"encoder.layers.{i}.", f"transformer.resblocks.{i}."
        sd[d + "attn.in_proj_weight"] = torch.cat([h[s + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
        sd[d + "attn.in_proj_bias"] = torch.cat([h[s + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
        for a, b in (("self_attn.out_proj", "attn.out_proj"), ("layer_norm1", "ln_1"), ("layer_norm2", "ln_2"),
                     ("mlp.fc1", "mlp.c_fc"), ("mlp.fc2", "mlp.c_proj")):
            sd[d + b + ".weight"], sd[d + b + ".bias"] = h[s + a + ".weight"], h[s + a + ".bias"]
    return m, {"model.visual." + k: t for k, t in sd.items()}


@torch.no_grad()
def clip_vision_forward(m, pixel_values):
    """(pooled embeds (B, E), ln_post tokens (B, 256, W)) as open_clip's VisionTransformer returns with output_tokens."""
    out = m(pixel_values=pixel_values)
    tokens = m.vision_model.post_layernorm(out.last_hidden_state[:, 1:])
    return out.image_embeds, tokens


def dinov2_reg(depth=2, width=256, seed=0, init_scale=3.0):
    """-> (HF model, dinov2-hub-style state dict with the `model.` prefix).  image_size 224 so that the position table
    already has the 16 x 16 grid (no interpolation on either side)."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    torch.manual_seed(seed)
    cfg = Dinov2WithRegistersConfig(hidden_size=width, num_hidden_layers=depth, num_attention_heads=width // 64, mlp_ratio=4,
                                    image_size=224, patch_size=14, num_register_tokens=4, layer_norm_eps=1e-6,
                                    hidden_act="gelu", initializer_range=0.02 * init_scale, layerscale_value=1.0)
    m = _randomize_norms(Dinov2WithRegistersModel(cfg).eval().float(), seed)
    h = m.state_dict()
    sd = {"cls_token": h["embeddings.cls_token"], "register_tokens": h["embeddings.register_tokens"],
          "pos_embed": h["embeddings.position_embeddings"], "patch_embed.proj.weight": h["embeddings.patch_embeddings.projection.weight"],
          "patch_embed.proj.bias": h["embeddings.patch_embeddings.projection.bias"],
          "norm.weight": h["layernorm.weight"], "norm.bias": h["layernorm.bias"]}
    for i in range(depth):
        s, d = f"encoder.layer.{i}.", f"blocks.{i}."
        sd[d + "attn.qkv.weight"] = torch.cat([h[s + f"attention.attention.{n}.weight"] for n in ("query", "key", "value")], 0)
        sd[d + "attn.qkv.bias"] = torch.cat([h[s + f"attention.attention.{n}.bias"] for n in ("query", "key", "value")], 0)
        for a, b in (("attention.output.dense", "attn.proj"), ("norm1", "norm1"), ("norm2", "norm2"), ("mlp.fc1", "mlp.fc1"),
                     ("mlp.fc2", "mlp.fc2")):
            sd[d + b + ".weight"], sd[d + b + ".bias"] = h[s + a + ".weight"], h[s + a + ".bias"]
        sd[d + "ls1.gamma"], sd[d + "ls2.gamma"] = h[s + "layer_scale1.lambda1"], h[s + "layer_scale2.lambda1"]
    return m, {"model." + k: t for k, t in sd.items()}
