"""Shared launch sequence of the PixArt-style denoisers (shared adaLN + per-block scale_shift_table, RMSNorm
pre-norms, T2IFinalLayer): `DiT_I23D_PixelArt` (dit/dit_i23d.py) and `DiT_TriLatent_PixelArt`
(dit/dit_trilatent.py) differ only in what conditions them -- I23D adds q/k RMSNorm inside both attentions
and the DINO tokens as a second self-attention K/V source."""
from __future__ import annotations

import torch

from .. import ops
from .._lib import NORM_NONE, NORM_RMS
from ._graph import ContextCache, ForwardGraph, capture_forward, graphs_enabled


def _split_residual_pass() -> bool:
    import os
    return os.environ.get("LN3_SPLIT_RESID_PASS", "1") != "0"


def pixart_forward(model, P: dict, cx: dict, ws: dict, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """x (B, 3C, S, S) fp32, t (B,) fp32; cx = cached step-invariant conditioning: 'cls' (B, D) fp32 pooled
    embedding added to t_emb, 'ckv' (L, B, Lc, 2D) cross-attention K|V, optional 'dkv' (L, B, Ld, 2D) DINO
    self-attention K|V, 'rows' / 'oconst' closed-form cross-attention of identical-token samples
    (DiT_TriLatent._context_kv).  Per-block weights in P['blocks'] carry 'qk_norm' / 'cq_norm' when the
    attentions are q/k-normalised."""
    B = x.shape[0]
    D, H, T = model.embed_dim, model.num_heads, model.pos_embed.shape[1]
    M = B * T
    ops.timestep_embedding(t, out=ws["tfeat"])
    ops.gemm(ws["tfeat"], P["t0_w"], P["t0_b"], act=ops.ACT_SILU, out=ws["th"])
    ws["t"].copy_(cx["cls"])                                                       # t = t_emb + pooled embedding
    ops.gemm(ws["th"], P["t2_w"], P["t2_b"], out_kind=ops.OUT_RESID_F32, out=ws["t"])
    ops.norm_modulate(ws["t"], norm=NORM_NONE, act=ops.ACT_SILU, out=ws["st"])
    ops.gemm(ws["st"], P["ada_w"], P["ada_b"], out_kind=ops.OUT_F32, out=ws["t0"])  # shared adaLN (B, 6D)
    torch.add(P["tables"][:, None, :], ws["t0"][None], out=ws["mod"])              # + per-block tables
    xs = ops.patch_embed(x, P["pe_w"], P["pe_b"], P["pos"], out=ws["x"])
    x2 = xs.view(M, D)
    qkv3, att3, q3 = ws["qkv"].view(B, T, 3 * D), ws["att"].view(B, T, D), ws["q"].view(B, T, D)
    (g0, g1), oconst = cx["rows"], cx["oconst"]
    r0, r1 = g0 * T, g1 * T
    val, pend_gate = ws["v"], None   # deferred residuals (see DiT_TriLatent._forward_impl)
    for l, W in enumerate(P["blocks"]):
        mod = ws["mod"][l]
        sl = lambda j: mod[:, j * D:(j + 1) * D]
        ops.norm_modulate(x2, norm=NORM_RMS, weight=W["n1_w"], eps=1e-5, shift=sl(0), scale=sl(1), mod_rows=T, out=ws["a"],
                          resid=val if l > 0 else None, resid_gate=pend_gate, resid_gate_rows=T)
        ops.gemm(ws["a"], W["qkv_w"], W["qkv_b"], out=ws["qkv"], head_norm=W.get("qk_norm"), head_norm_sec_cols=D)
        if "dkv" in cx:
            dkv = cx["dkv"][l]
            ops.fmha(qkv3[:, :, :D], qkv3[:, :, D:2 * D], qkv3[:, :, 2 * D:], H, out=att3, k2=dkv[:, :, :D], v2=dkv[:, :, D:])
        else:
            ops.fmha(qkv3[:, :, :D], qkv3[:, :, D:2 * D], qkv3[:, :, 2 * D:], H, out=att3)
        ops.gemm(ws["att"], W["proj_w"], W["proj_b"], out=val)
        split = oconst is not None and _split_residual_pass()      # see DiT_TriLatent._forward_impl
        if split:
            if r1 > r0:
                ops.norm_modulate(x2[r0:r1], norm=NORM_NONE, out=ws["xb"][r0:r1], resid=val[r0:r1],
                                  resid_gate=sl(2)[g0:g1], resid_gate_rows=T)
        else:
            ops.norm_modulate(x2, norm=NORM_NONE, out=ws["xb"], resid=val, resid_gate=sl(2), resid_gate_rows=T)
        if r1 > r0:
            ops.gemm(ws["xb"][r0:r1], W["cq_w"], out=ws["q"][r0:r1], head_norm=W.get("cq_norm"), head_norm_sec_cols=D)
            ckv = cx["ckv"][l]
            ops.fmha(q3[g0:g1], ckv[g0:g1, :, :D], ckv[g0:g1, :, D:], H, out=att3[g0:g1])
            ops.gemm(ws["att"][r0:r1], W["co_w"], W["co_b"], out=val[r0:r1])
        ops.norm_modulate(x2, norm=NORM_RMS, weight=W["n2_w"], eps=1e-5, shift=sl(3), scale=sl(4), mod_rows=T, out=ws["a"],
                          resid=val, resid_bcast=oconst[l] if oconst is not None else None, resid_bcast_rows=T,
                          resid_rows=(r0, r1) if oconst is not None else None,
                          resid_out_gate=sl(2) if split else None, resid_out_gate_rows=T)
        ops.gemm(ws["a"], W["fc1_w"], W["fc1_b"], act=ops.ACT_GELU_ERF, out=ws["h"])
        ops.gemm(ws["h"], W["fc2_w"], W["fc2_b"], out=val)
        pend_gate = sl(5)
    ops.norm_modulate(x2, norm=NORM_NONE, resid=val, resid_gate=pend_gate, resid_gate_rows=T, want_out=False)
    # T2IFinalLayer: shift = table[0] + t, scale = table[1] + t
    return ops.final_layer(xs, ws["t"], ws["t"], P["fin_w"], P["fin_b"], model.input_size,
                           shift_tab=P["fin_tab"][0].contiguous(), scale_tab=P["fin_tab"][1].contiguous())


def pixart_workspace(model, B: int) -> dict:
    dev = model.pos_embed.device
    D, T = model.embed_dim, model.pos_embed.shape[1]
    M = B * T
    e = lambda *s, dt=torch.bfloat16: torch.empty(*s, device=dev, dtype=dt)
    return dict(tfeat=e(B, 256), th=e(B, D), t=e(B, D, dt=torch.float32), st=e(B, D),
                t0=e(B, 6 * D, dt=torch.float32), mod=e(model.depth, B, 6 * D, dt=torch.float32),
                x=e(B, T, D, dt=torch.float32), xb=e(M, D), a=e(M, D), v=e(M, D), qkv=e(M, 3 * D), att=e(M, D), q=e(M, D),
                h=e(M, int(model.mlp_ratio) * D))


class PixArtGraphMixin:
    """Derived-state management shared by the PixArt-style denoisers: bf16 repacks (`_prep`), per-batch
    workspaces, the step-invariant conditioning cache with its model-owned static buffers, and the
    CUDA-graph cache of a forward (see dit/_graph.py)."""

    def _invalidate(self):
        self._prep = None
        self._ws = {}
        self._ctx_cache = ContextCache()
        self._ctx_static = {}
        self._graphs = {}

    def _apply(self, fn, *a, **kw):
        self._invalidate()
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._invalidate()
        return super().load_state_dict(*a, **kw)

    def _static(self, key, make):
        st = self._ctx_static.get(key)
        if st is None:
            st = self._ctx_static[key] = make()
        return st

    def _workspace(self, B):
        ws = self._ws.get(B)
        if ws is None:
            ws = self._ws[B] = pixart_workspace(self, B)
        return ws

    def _graph(self, B, cx) -> ForwardGraph:
        key = (B, cx["ckv"].shape[2], cx["dkv"].shape[2] if "dkv" in cx else 0, cx["rows"], cx["oconst"] is not None)
        g = self._graphs.get(key)
        if g is None:
            dev = self.pos_embed.device
            g = ForwardGraph()
            g.key, g.cross_attention_rows = key, cx["rows"]
            g.x = torch.zeros(B, 3 * self.in_channels, self.input_size, self.input_size, device=dev)
            g.t = torch.zeros(B, device=dev)
            ws = self._workspace(B)
            capture_forward(g, lambda: pixart_forward(self, self._prep, cx, ws, g.x, g.t), dev)
            self._graphs[key] = g
        return g

    def _run(self, x, t, cx):
        """One forward: a replay of the cached CUDA graph (captured on first use), or the eager launch
        sequence under LN3_CUDA_GRAPH=0 / inside a caller's own capture."""
        if graphs_enabled() and not torch.cuda.is_current_stream_capturing():
            g = self._graph(x.shape[0], cx)
            g.x.copy_(x)
            g.t.copy_(t)
            g.replay()
            return g.out.clone()
        return pixart_forward(self, self._prep, cx, self._workspace(x.shape[0]), x.float().contiguous(), t)
