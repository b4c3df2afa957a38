"""TEST INFRASTRUCTURE ONLY -- stand-ins for third-party packages the reference imports but
that are absent from this image, so that the reference's own hot-path modules can be imported
from /root/reference *in the build container* to validate the oracle restatement and to generate
tests/golden/* (see oracle/make_golden.py).  Nothing here is imported by the product package.

Semantics assumed for the un-vendored arithmetic ("parity unpinned" for
these third-party pieces):
  xformers 0.0.26 memory_efficient_attention = softmax(q k^T / sqrt(d)) v, no mask, for both the
      (B, M, H, K) and (B*H, M, K) layouts;
  xformers FusedMLP at eval / p=0 = Linear(no bias) -> + bias -> exact-erf GELU -> Linear(no bias)
      -> + bias with parameter names mlp.0.weight, mlp.1.bias, mlp.2.weight, mlp.3.bias;
  timm 0.6.13 PatchEmbed = Conv2d(k = s = patch) -> flatten(2).transpose(1, 2); Mlp = fc1-act-fc2;
  torchdiffeq 0.2.3 odeint: only the fixed-grid 'euler' / 'heun' / 'midpoint' solvers.
"""
from __future__ import annotations

import math
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ----------------------------------------------------------------------------- xformers
def memory_efficient_attention(q, k, v, attn_bias=None, op=None, p=0.0, scale=None):
    assert attn_bias is None
    if q.dim() == 4:  # (B, M, H, K)
        q_, k_, v_ = (t.permute(0, 2, 1, 3) for t in (q, k, v))
        s = (q_ @ k_.transpose(-1, -2)) * (scale or q.shape[-1] ** -0.5)
        return (s.float().softmax(-1).to(q.dtype) @ v_).permute(0, 2, 1, 3)
    s = (q @ k.transpose(-1, -2)) * (scale or q.shape[-1] ** -0.5)  # (B*H, M, K)
    return s.float().softmax(-1).to(q.dtype) @ v


class _BiasAdd(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(n))

    def forward(self, x):
        return x + self.bias


class _GeluThenBias(nn.Module):
    """xformers' fused dropout(+bias)+activation layer at p = 0: activation(x + bias)."""

    def __init__(self, n):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(n))

    def forward(self, x):
        return F.gelu(x + self.bias)


class FusedMLP(nn.Module):
    def __init__(self, dim_model, dropout, activation, hidden_layer_multiplier, bias=True, **kw):
        super().__init__()
        assert dropout == 0
        hid = hidden_layer_multiplier * dim_model
        self.mlp = nn.Sequential(
            nn.Linear(dim_model, hid, bias=False),
            _GeluThenBias(hid),
            nn.Linear(hid, dim_model, bias=False),
            _BiasAdd(dim_model),
        )

    def forward(self, x):
        return self.mlp(x)


class _Activation:
    GeLU = "gelu"


# ----------------------------------------------------------------------------- timm
class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True, bias=True):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size,
                              bias=bias)
        self.norm = nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 bias=True, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer() if isinstance(act_layer, type) or callable(act_layer) and not isinstance(
            act_layer, nn.Module) else act_layer
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


# ----------------------------------------------------------------------------- torchdiffeq
def odeint(fn, x, t, method="euler", atol=None, rtol=None, **kw):
    """Fixed-grid solvers of torchdiffeq 0.2.3 (no sub-stepping: step_size = grid spacing)."""
    ys = [x]
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        y = ys[-1]
        if method == "euler":
            y = y + dt * fn(t0, y)
        elif method == "midpoint":
            half = 0.5 * dt
            y = y + dt * fn(t0 + half, y + half * fn(t0, y))
        elif method == "heun":  # torchdiffeq's "heun2"/"heun3" differ; transport uses 'heun'
            k1 = fn(t0, y)
            k2 = fn(t1, y + dt * k1)
            y = y + 0.5 * dt * (k1 + k2)
        else:
            raise NotImplementedError(f"adaptive / unknown solver '{method}' is not stubbed")
        ys.append(y)
    return torch.stack(ys, 0)


# ----------------------------------------------------------------------------- install
def install(reference_root: str = REFERENCE_ROOT) -> None:
    """Register the stubs and make the reference's hot-path sub-modules importable by path."""
    if "xformers" in sys.modules and getattr(sys.modules["xformers"], "_ln3_stub", False):
        return
    xf = _mod("xformers", _ln3_stub=True)
    xops = _mod("xformers.ops", memory_efficient_attention=memory_efficient_attention,
                unbind=torch.unbind, MemoryEfficientAttentionFlashAttentionOp=None)
    xops.fmha = _mod("xformers.ops.fmha", MemoryEfficientAttentionFlashAttentionOp=None)
    xf.ops = xops
    comp = _mod("xformers.components")
    act = _mod("xformers.components.activations", Activation=_Activation,
               build_activation=lambda a: nn.GELU())
    ff = _mod("xformers.components.feedforward")
    fm = _mod("xformers.components.feedforward.fused_mlp", FusedMLP=FusedMLP)
    ff.fused_mlp = fm
    comp.activations, comp.feedforward = act, ff
    xf.components = comp
    xf.triton = _mod("xformers.triton", FusedLayerNorm=nn.LayerNorm)

    timm = _mod("timm")
    tm = _mod("timm.models")
    tv = _mod("timm.models.vision_transformer", PatchEmbed=PatchEmbed, Mlp=Mlp, Attention=nn.Module)
    tl = _mod("timm.models.layers", DropPath=nn.Identity, to_2tuple=lambda x: (x, x),
              trunc_normal_=nn.init.trunc_normal_)
    timm.models, tm.vision_transformer, tm.layers = tm, tv, tl

    _mod("torchdiffeq", odeint=odeint)

    class _Cfg(dict):
        pass

    _mod("omegaconf", ListConfig=list, OmegaConf=_Cfg, DictConfig=dict)
    _mod("blobfile", BlobFile=open, join=lambda *a: "/".join(a))

    # Namespace packages that bypass the trainer-heavy package __init__ files.
    for pkg in ("nsr", "guided_diffusion", "sgm", "sgm.modules", "sgm.modules.diffusionmodules",
                "dit", "vit", "ldm", "ldm.modules", "ldm.modules.diffusionmodules", "transport_ns"):
        if pkg == "transport_ns":
            continue
        m = types.ModuleType(pkg)
        m.__path__ = [f"{reference_root}/{pkg.replace('.', '/')}"]
        sys.modules[pkg] = m
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


def patch_dit_namespace() -> None:
    """dit_models_xformers only imports fused_mlp/Activation when CUDA is available
    (dit/dit_models_xformers.py:39-43); inject them for the CPU oracle."""
    import dit.dit_models_xformers as dmx  # noqa
    dmx.fused_mlp = sys.modules["xformers.components.feedforward.fused_mlp"]
    dmx.Activation = _Activation
    import vit.vision_transformer as vvt  # noqa
    # MemEffAttention silently drops q/k-norm without xformers (vision_transformer.py:109-111):
    # the oracle must take the xformers branch.
    vvt.XFORMERS_AVAILABLE = True
    vvt.memory_efficient_attention = memory_efficient_attention
    vvt.unbind = torch.unbind
