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


def attention(q, k, v, split=0, out0=None, out1=None):
    """q,k,v bf16 [B,H,S,hd] (hd 64 or 128) -> (out0 [B*split, H*hd] or None, out1 [B*(S-split), H*hd])."""
    B, H, S, hd = q.shape
    assert hd in (64, 128) and q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    if split > 0 and out0 is None:
        out0 = torch.empty((B * split, H * hd), dtype=torch.bfloat16, device=q.device)
    if split < S and out1 is None:
        out1 = torch.empty((B * (S - split), H * hd), dtype=torch.bfloat16, device=q.device)
    _lib.check(_lib.load().drag_attention_bf16(
        _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), B, H, S, hd, split, _lib.ptr(out0), out0.stride(0) if out0 is not None else 8,
        _lib.ptr(out1), out1.stride(0) if out1 is not None else 8, _lib.current_stream_ptr(q.device)),
        "drag_attention_bf16")
    return out0, out1


def layernorm(x, mul=None, add=None, adaln=False, rows_per_batch=0, eps=1e-6, out=None):
    """LayerNorm without affine followed by * (1 + mul[b]) + add[b] (adaln) or * mul + add (affine)."""
    M, d = x.shape
    if out is None:
        out = torch.empty((M, d), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().drag_layernorm_bf16(
        _lib.ptr(x), x.stride(0), _lib.ptr(out), out.stride(0), M, d, _lib.ptr(mul),
        (mul.stride(0) if (mul is not None and mul.dim() == 2) else 0), _lib.ptr(add),
        (add.stride(0) if (add is not None and add.dim() == 2) else 0), int(adaln), rows_per_batch, eps,
        _lib.current_stream_ptr(x.device)), "drag_layernorm_bf16")
    return out


def timestep_embed(t):
    out = torch.empty((t.shape[0], 256), dtype=torch.bfloat16, device=t.device)
    _lib.check(_lib.load().drag_timestep_embed(_lib.ptr(t), _lib.ptr(out), t.shape[0],
                                               _lib.current_stream_ptr(t.device)), "drag_timestep_embed")
    return out


def l2_normalize(x):
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.load().drag_l2_normalize(_lib.ptr(x), _lib.ptr(out), x.shape[0], x.shape[1],
                                             _lib.current_stream_ptr(x.device)), "drag_l2_normalize")
    return out


def debug_set(key: int, value: int) -> None:
    _lib.check(_lib.load().drag_debug_set(key, value), "drag_debug_set")
