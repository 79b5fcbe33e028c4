"""Thin torch-tensor wrappers over the C-ABI compute entry points (device pointers + current stream).
No arithmetic happens here and nothing falls back to PyTorch."""
from __future__ import annotations

import torch

from . import _lib

EPI_BIAS, EPI_GELU_TANH, EPI_QUICK_GELU, EPI_SILU, EPI_GATE_RESID, EPI_QKV_ROPE, EPI_BIAS_F32 = range(7)


def _bf16(t, name):
    if t is not None and not (t.is_cuda and t.dtype == torch.bfloat16):
        raise TypeError(f"{name}: need a CUDA bfloat16 tensor")
    return t


def linear(a, w, bias=None, mode=EPI_BIAS, out=None, resid=None, gate=None, rows_per_batch=0):
    """out = epilogue(a[M,K] @ w[N,K]^T + bias) on the tcgen05 GEMM. a/w may be row-strided views."""
    _bf16(a, "a"); _bf16(w, "w"); _bf16(bias, "bias"); _bf16(resid, "resid"); _bf16(gate, "gate")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), device=a.device,
                          dtype=torch.float32 if mode == EPI_BIAS_F32 else torch.bfloat16)
    assert out.stride(1) == 1
    _lib.check(_lib.load().drag_gemm_bf16(
        _lib.ptr(a), a.stride(0), _lib.ptr(w), w.stride(0), M, N, K, mode, _lib.ptr(bias), _lib.ptr(out),
        out.stride(0), _lib.ptr(resid), resid.stride(0) if resid is not None else 0, _lib.ptr(gate),
        gate.stride(0) if gate is not None else 0, rows_per_batch, _lib.current_stream_ptr(a.device)),
        "drag_gemm_bf16")
    return out


def qkv_rope(a, w, bias, q_out, k_out, v_out, q_norm_w, k_norm_w, rope_cos, rope_sin, tok_offset,
             rows_per_batch, eps=1e-6):
    """Fused QKV projection + per-head RMSNorm + RoPE, scattered into [B,H,S_total,128] buffers."""
    M, K = a.shape
    B, H, S, hd = q_out.shape
    assert hd == 128 and w.shape[0] == 3 * H * 128
    _lib.check(_lib.load().drag_gemm_qkv_rope(
        _lib.ptr(a), a.stride(0), _lib.ptr(w), w.stride(0), M, K, H, _lib.ptr(bias), _lib.ptr(q_out),
        _lib.ptr(k_out), _lib.ptr(v_out), _lib.ptr(q_norm_w), _lib.ptr(k_norm_w), _lib.ptr(rope_cos),
        _lib.ptr(rope_sin), S, tok_offset, rows_per_batch, eps, _lib.current_stream_ptr(a.device)),
        "drag_gemm_qkv_rope")
