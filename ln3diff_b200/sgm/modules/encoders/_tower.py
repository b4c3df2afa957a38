"""Pre-LayerNorm transformer tower on the libln3b200 kernels: the engine behind the three frozen conditioner
encoders (CLIP text, OpenCLIP ViT-L/14 image, DINOv2 ViT-L/14-reg).

One layer (transformers CLIPEncoderLayer / open_clip ResidualAttentionBlock / dinov2 NestedTensorBlock at eval):
    x = x + ls1 * proj(attn(LN1(x)))          attn: softmax(q k^T / 8 [+ causal mask]) v per 64-wide head
    x = x + ls2 * fc2(act(fc1(LN2(x))))       act: QuickGELU (CLIP) or erf-GELU (DINOv2); ls = LayerScale or 1
Kernel sequence per layer (same pattern as the DiT blocks): ln3_norm_modulate applies the PREVIOUS projection's
deferred residual and emits the normalised bf16 GEMM operand; ln3_gemm_bf16 (bias / activation epilogues);
ln3_fmha_fwd on strided views of the packed qkv buffer (no permutes).  The residual stream is fp32.
"""
from __future__ import annotations

import torch

from .... import ops
from ...._lib import NORM_LAYER, NORM_NONE


class TowerLayer:
    __slots__ = ("ln1", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1", "ln2", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2")


def _ln_pair(gamma: torch.Tensor, beta: torch.Tensor, device):
    """LayerNorm affine as the (shift, scale) pair of ln3_norm_modulate: y = LN(x) * (1 + (gamma - 1)) + beta."""
    g = gamma.detach().to(device=device, dtype=torch.float32)
    b = beta.detach().to(device=device, dtype=torch.float32)
    return (b.reshape(1, -1).contiguous(), (g - 1.0).reshape(1, -1).contiguous())


def _w(t: torch.Tensor, device):
    return t.detach().to(device=device, dtype=torch.bfloat16).contiguous()


def _b(t: torch.Tensor, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class TransformerTower:
    """Frozen weights repacked once (bf16 GEMM operands, fp32 biases / LayerNorm / LayerScale vectors)."""

    def __init__(self, width: int, heads: int, act: int, eps: float, causal: bool, device):
        assert width % 128 == 0 and width // heads == 64, "kernels: width % 128 == 0, head_dim 64"
        self.width, self.heads, self.act, self.eps, self.causal = width, heads, act, eps, causal
        self.device = torch.device(device)
        self.layers: list[TowerLayer] = []

    def add_layer(self, *, ln1, qkv_w, qkv_b, proj_w, proj_b, ln2, fc1_w, fc1_b, fc2_w, fc2_b, ls1=None, ls2=None):
        d = self.device
        L = TowerLayer()
        L.ln1, L.ln2 = _ln_pair(*ln1, d), _ln_pair(*ln2, d)
        L.qkv_w, L.qkv_b = _w(qkv_w, d), _b(qkv_b, d)
        L.proj_w, L.proj_b = _w(proj_w, d), _b(proj_b, d)
        L.fc1_w, L.fc1_b, L.fc2_w, L.fc2_b = _w(fc1_w, d), _b(fc1_b, d), _w(fc2_w, d), _b(fc2_b, d)
        L.ls1 = None if ls1 is None else _b(ls1, d).reshape(1, -1)
        L.ls2 = None if ls2 is None else _b(ls2, d).reshape(1, -1)
        self.layers.append(L)

    @torch.no_grad()
    def run(self, x: torch.Tensor, n_layers: int | None = None, hidden_states: list | None = None):
        """x fp32 (B, L, D) contiguous: the embedded input, updated IN PLACE into the residual stream after the
        last requested layer.  hidden_states (a list) receives a copy of the stream before every layer and after
        the last one (transformers' `output_hidden_states`)."""
        B, Lq, D = x.shape
        assert D == self.width and x.is_contiguous() and x.dtype == torch.float32
        rows = B * Lq
        xs = x.view(rows, D)
        pend = None   # (bf16 value, LayerScale gate or None) of the projection whose residual is not applied yet
        layers = self.layers if n_layers is None else self.layers[:n_layers]

        def norm(pair, pend_):
            kw = {}
            if pend_ is not None:
                kw = dict(resid=pend_[0], resid_gate=pend_[1], resid_gate_rows=rows)
            return ops.norm_modulate(xs, norm=NORM_LAYER, shift=pair[0], scale=pair[1], mod_rows=rows, eps=self.eps, **kw)

        for L in layers:
            if hidden_states is not None:
                if pend is not None:   # materialise the stream: apply the deferred residual now
                    ops.norm_modulate(xs, norm=NORM_NONE, resid=pend[0], resid_gate=pend[1], resid_gate_rows=rows,
                                      want_out=False)
                    pend = None
                hidden_states.append(x.clone())
            h = norm(L.ln1, pend)
            qkv = ops.gemm(h, L.qkv_w, L.qkv_b).view(B, Lq, 3 * D)
            o = ops.fmha(qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:], self.heads, causal=self.causal)
            val = ops.gemm(o.view(rows, D), L.proj_w, L.proj_b)
            h = norm(L.ln2, (val, L.ls1))
            f = ops.gemm(h, L.fc1_w, L.fc1_b, act=self.act)
            val = ops.gemm(f, L.fc2_w, L.fc2_b)
            pend = (val, L.ls2)
        if pend is not None:
            ops.norm_modulate(xs, norm=NORM_NONE, resid=pend[0], resid_gate=pend[1], resid_gate_rows=rows, want_out=False)
        if hidden_states is not None:
            hidden_states.append(x.clone())
        return x

    @torch.no_grad()
    def layer_norm(self, x: torch.Tensor, gamma_beta_pair) -> torch.Tensor:
        """LayerNorm with affine over the last dim of fp32 (..., D) -> fp32 (bf16-rounded: the kernel emits the GEMM
        operand type; every consumer of a conditioner output feeds it to a bf16 GEMM)."""
        D = x.shape[-1]
        xs = x.reshape(-1, D).contiguous()
        out = ops.norm_modulate(xs, norm=NORM_LAYER, shift=gamma_beta_pair[0], scale=gamma_beta_pair[1],
                                mod_rows=xs.shape[0], eps=self.eps)
        return out.float().view(*x.shape)
