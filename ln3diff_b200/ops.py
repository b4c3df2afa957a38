"""Thin, checked Python entry points over the C ABI (one function per exported op).

Shape / dtype / contiguity violations raise here (mirroring the TORCH_CHECKs of the reference's
own native ops, utils/torch_utils/ops/bias_act.cpp:39-60); everything else is the C call.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, ACT_SILU, OUT_BF16, OUT_F32,
                   OUT_RESID_F32)

__all__ = ["gemm", "ACT_NONE", "ACT_GELU_ERF", "ACT_GELU_TANH", "ACT_SILU", "OUT_BF16", "OUT_F32",
           "OUT_RESID_F32"]


def _req(cond: bool, msg: str) -> None:
    if not cond:
        raise ValueError(msg)


def _cuda(t: torch.Tensor, name: str, dtype=None) -> None:
    _req(t.is_cuda, f"{name} must be a CUDA tensor (no CPU fallback)")
    if dtype is not None:
        _req(t.dtype == dtype, f"{name} must be {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *,
         act: int = ACT_NONE, out_kind: int = OUT_BF16, out: torch.Tensor | None = None,
         out2: torch.Tensor | None = None, gate: torch.Tensor | None = None,
         gate_rows: int = 1) -> torch.Tensor:
    """out = epilogue(a @ w.T); a (M,K) bf16, w (N,K) bf16 (nn.Linear weight layout).

    OUT_RESID_F32: `out` is the fp32 residual stream (M,N), updated in place with
    out += gate[(m // gate_rows)] * val; `gate` is a 2-D fp32 view (G,N) with unit inner stride.
    """
    _cuda(a, "a", torch.bfloat16)
    _cuda(w, "w", torch.bfloat16)
    _req(a.dim() == 2 and w.dim() == 2, "a and w must be 2-D")
    _req(a.stride(1) == 1 and w.stride(1) == 1, "a and w must have unit inner stride")
    M, K = a.shape
    N, K2 = w.shape
    _req(K == K2, f"inner dims differ: {K} vs {K2}")
    if out is None:
        _req(out_kind != OUT_RESID_F32, "OUT_RESID_F32 needs the residual tensor as `out`")
        out = torch.empty((M, N), device=a.device,
                          dtype=torch.bfloat16 if out_kind == OUT_BF16 else torch.float32)
    _cuda(out, "out", torch.bfloat16 if out_kind == OUT_BF16 else torch.float32)
    _req(out.shape == (M, N) and out.stride(1) == 1, "out must be (M,N) with unit inner stride")
    args = _lib.GemmArgs()
    args.A, args.W, args.out = a.data_ptr(), w.data_ptr(), out.data_ptr()
    args.M, args.N, args.K = M, N, K
    args.lda, args.ldw, args.ldo = a.stride(0), w.stride(0), out.stride(0)
    if bias is not None:
        _cuda(bias, "bias", torch.float32)
        _req(bias.shape == (N,) and bias.is_contiguous(), "bias must be contiguous (N,)")
        args.bias = bias.data_ptr()
    if out2 is not None:
        _cuda(out2, "out2", torch.bfloat16)
        _req(out2.shape == (M, N) and out2.stride(1) == 1, "out2 must be (M,N)")
        args.out2, args.ldo2 = out2.data_ptr(), out2.stride(0)
    if gate is not None:
        _cuda(gate, "gate", torch.float32)
        _req(gate.dim() == 2 and gate.shape[1] == N and gate.stride(1) == 1, "gate must be (G,N)")
        _req(gate.shape[0] * gate_rows >= M, "gate has too few rows")
        args.gate, args.gate_ld, args.gate_rows = gate.data_ptr(), gate.stride(0), gate_rows
    args.act, args.out_kind = act, out_kind
    _lib.check(_lib.lib().ln3_gemm_bf16(C.byref(args), _lib.current_stream()), "ln3_gemm_bf16")
    return out
