"""GPU: the conditioner towers (reference sgm/modules/encoders/modules.py:80-190,347-408,578-868) against
the fp32 transformers models on the CPU (oracle/conditioners.py: the reference's own dependency for the text tower,
transformers' ports of the same architectures for the two image towers)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a GPU"
    from ln3diff_b200 import _lib
    _lib.lib()
    return torch.device("cuda", 0)


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_fmha_causal_mask(dev):
    from ln3diff_b200 import ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    for (B, H, L) in [(3, 12, 77), (2, 4, 200), (1, 2, 1)]:
        D = H * 64
        qkv = torch.randn(B, L, 3 * D, generator=g).bfloat16()
        d = qkv.to(dev)
        out = ops.fmha(d[:, :, :D], d[:, :, D:2 * D], d[:, :, 2 * D:], H, causal=True)
        qf, kf, vf = (t.float().reshape(B, L, H, 64).transpose(1, 2) for t in (qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]))
        ref = F.scaled_dot_product_attention(qf, kf, vf, is_causal=True).transpose(1, 2).reshape(B, L, D)
        assert _rel(out, ref) < 6e-3
    with pytest.raises(ValueError, match="single K/V source"):
        ops.fmha(d[:, :, :D], d[:, :, D:2 * D], d[:, :, 2 * D:], H, k2=d[:, :, D:2 * D], v2=d[:, :, 2 * D:], causal=True)


def test_gemm_quick_gelu(dev):
    from ln3diff_b200 import ops
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(300, 128, generator=g) * 0.5).bfloat16()
    w = (torch.randn(256, 128, generator=g) * 0.2).bfloat16()
    b = torch.randn(256, generator=g)
    lin = a.float() @ w.float().t() + b
    out = ops.gemm(a.to(dev), w.to(dev), b.to(dev), act=ops.ACT_QUICK_GELU)
    assert _rel(out, lin * torch.sigmoid(1.702 * lin)) < 4e-3


def test_clip_text_tower_matches_transformers(dev):
    """Full-size CLIP-L text tower (12 x 768, 77 tokens, causal, QuickGELU): last hidden state, pooled output at the
    EOS position, an intermediate hidden state; state dict in the reference object's layout (`transformer.*`)."""
    from ln3diff_b200.sgm.modules.encoders.modules import FrozenCLIPEmbedder
    from oracle import conditioners as oc
    hf, sd = oc.clip_text(depth=12)
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(3, 49000, (4, 77), generator=g)
    for b, pos in enumerate((5, 20, 76, 40)):
        ids[b, pos] = 49407          # EOS = largest id (legacy eos_token_id 2 -> argmax)
        ids[b, pos + 1:] = 49407
    with torch.no_grad():
        ref = hf(input_ids=ids, output_hidden_states=True)
    emb = FrozenCLIPEmbedder(device=dev, always_return_pooled=True, state_dict=sd)
    z, pooled = emb(ids)
    assert z.shape == (4, 77, 768) and pooled.shape == (4, 768) and z.dtype == torch.float32
    assert _rel(z, ref.last_hidden_state) < 1e-2
    assert _rel(pooled, ref.pooler_output) < 1e-2
    hid = FrozenCLIPEmbedder(device=dev, layer="hidden", layer_idx=-2, state_dict=sd)(ids)
    assert _rel(hid, ref.hidden_states[-2]) < 1e-2
    # a second, different batch through the same object (no stale state)
    ids2 = ids.flip(0)
    z2, _ = emb(ids2)
    assert torch.equal(z2.cpu(), z.flip(0).cpu())
    with pytest.raises(RuntimeError, match="tokenizer"):
        emb(["a chair"])


def test_openclip_image_tower_matches_transformers_port(dev):
    from ln3diff_b200.sgm.modules.encoders.modules import FrozenOpenCLIPImageEmbedder
    from oracle import conditioners as oc
    hf, sd = oc.clip_vision(depth=3, width=256, mlp=1024, embed=128)
    emb = FrozenOpenCLIPImageEmbedder(device=dev, output_tokens=True, state_dict=sd)
    g = torch.Generator().manual_seed(2)
    img = torch.rand(3, 3, 224, 224, generator=g) * 2 - 1            # the reference feeds [-1, 1] images
    tokens, z = emb(img.to(dev))
    pix = emb.preprocess(img.to(dev)).cpu()                          # 224 input: the resize is the identity
    ref_z, ref_tokens = oc.clip_vision_forward(hf, pix)
    assert tokens.shape == (3, 256, 256) and z.shape == (3, 128)
    assert _rel(tokens, ref_tokens) < 1.5e-2 and _rel(z, ref_z) < 1.5e-2
    # pooled-only mode and the GeneralConditioner key routing
    emb2 = FrozenOpenCLIPImageEmbedder(device=dev, state_dict=sd)
    assert _rel(emb2(img.to(dev)), ref_z) < 1.5e-2


def test_dinov2_tower_matches_transformers_port_and_conditioner_concat(dev):
    from ln3diff_b200.sgm.modules.encoders.modules import (FrozenDinov2ImageEmbedder, FrozenOpenCLIPImageEmbedder,
                                                           GeneralConditioner)
    from oracle import conditioners as oc
    hf, sd = oc.dinov2_reg(depth=3, width=256)
    emb = FrozenDinov2ImageEmbedder(device=dev, state_dict=sd)
    g = torch.Generator().manual_seed(4)
    img = torch.rand(2, 3, 224, 224, generator=g) * 2 - 1
    tokens = emb(img.to(dev))
    pix = emb.preprocess(img.to(dev)).cpu()
    with torch.no_grad():
        ref = hf(pixel_values=pix).last_hidden_state
    assert tokens.shape == (2, 256, 256)
    assert _rel(tokens, ref[:, 5:]) < 1.5e-2
    feats = emb.forward_features(emb.preprocess(img.to(dev)))
    assert _rel(feats["x_norm_clstoken"], ref[:, 0]) < 1.5e-2
    # the I23D conditioner (sgm/configs/img23d-clipl-compat-fm-lognorm.yaml:22-43): CLIP tokens || DINO tokens on dim 2
    _, sdc = oc.clip_vision(depth=1, width=256, mlp=1024, embed=128)
    clip = FrozenOpenCLIPImageEmbedder(device=dev, output_tokens=True, state_dict=sdc)
    clip._emb_config = {"input_key": "img", "ucg_rate": 0.1}
    emb._emb_config = {"input_key": "img", "ucg_rate": 0.1}
    cond = GeneralConditioner([clip, emb])
    c, uc = cond.get_unconditional_conditioning({"img": img.to(dev)}, force_uc_zero_embeddings=["img"])
    assert c["crossattn"].shape == (2, 256, 512) and c["vector"].shape == (2, 128)
    assert torch.equal(c["crossattn"][:, :, 256:], tokens) and float(uc["crossattn"].abs().max()) == 0.0
    assert clip.ucg_rate == 0.1      # restored after the forced no-drop pass


def test_image_preprocess_downscale(dev):
    """kornia-style antialiased bicubic resize: constants are preserved, output range sane, 512 -> 224."""
    from ln3diff_b200.sgm.modules.encoders.modules import kornia_resize_bicubic
    x = torch.full((1, 3, 512, 512), 0.25, device=dev)
    y = kornia_resize_bicubic(x, (224, 224), antialias=True)
    assert y.shape == (1, 3, 224, 224) and float((y - 0.25).abs().max()) < 1e-5
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 448, 448, generator=g).to(dev)
    y = kornia_resize_bicubic(x, (224, 224), antialias=True)
    assert abs(float(y.mean()) - float(x.mean())) < 5e-3 and float(y.std()) < float(x.std())


def test_text_to_3d_through_the_conditioner(dev):
    """eval_cldm for one caption: conditioner -> sampler -> decode -> render; the conditioner's output feeds the DiT
    exactly as a hand-built context does (same latents), and the unconditional half is the zero embedding."""
    from ln3diff_b200 import pipeline
    from ln3diff_b200.sgm.modules.encoders.modules import FrozenCLIPEmbedder, GeneralConditioner
    from ln3diff_b200.utils import build_ae_decoder, build_t23d, orbit_cameras
    emb = FrozenCLIPEmbedder(device=dev, depth=2, seed=3, random_init=True)
    emb._emb_config = {"input_key": "caption", "ucg_rate": 0.1}
    cond = GeneralConditioner([emb])
    ids = torch.randint(3, 49000, (1, 77), generator=torch.Generator().manual_seed(1))
    ids[0, 9:] = 49407
    c, uc = pipeline.condition_prompt(cond, "caption", ids, num_samples=2, device=dev)
    assert c["crossattn"].shape == (2, 77, 768) and float(uc["crossattn"].abs().max()) == 0.0
    assert torch.equal(c["crossattn"][0], c["crossattn"][1])
    m = build_t23d("DiT-B/2", device=dev)
    dec = build_ae_decoder("DiT2-S/2", device=dev)
    cams = orbit_cameras(2).to(dev)
    lat, out = pipeline.text_to_3d(cond, m, dec, ids, cams, num_samples=2, num_steps=3, resolution=32)
    g = torch.Generator().manual_seed(41)
    randn = torch.randn(2, 12, 32, 32, generator=g).to(dev)
    lat2 = pipeline.sample_t23d(m, randn, c, uc, 3, 6.5)
    assert torch.equal(lat, lat2) and bool(torch.isfinite(out["image_raw"]).all())
    assert out["image_raw"].shape[-2:] == (32, 32)
