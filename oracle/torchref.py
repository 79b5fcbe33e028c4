"""torch-eager restatement of the hot path ON THE GPU (`--backend torch-ref`).
TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's `gpu_baseline` leg, never by domain_rag_b200/.

The reference's arithmetic is diffusers / torch eager kernels (cuBLAS GEMMs, SDPA, unfused LayerNorm / RoPE /
elementwise) at batch 1 in bf16 (batch_generate_flux_kshot.py:467-474, outpainting_updown_sampling_redux.py:1246-1257).
diffusers is not installable offline, so "the reference's torch path on the same box" is the oracle's own modules
(oracle/flux.py, oracle/vae.py - op for op what diffusers runs) executed on the B200:

  * dtype=torch.bfloat16 : what the reference runs (bf16 weights and activations, fp32 inside LayerNorm / RMSNorm /
    RoPE / softmax like torch does) -> the same-box GPU baseline and the bf16 noise floor of the parity tests;
  * dtype=torch.float32  : the oracle at FULL depth and width (19 + 38 blocks, 11.9 B parameters, S = 5337) - the CPU
    cannot follow this (88.85 TFLOP per forward), the B200 does it in seconds with TF32 disabled.

Weights are shared with the product path by reference (a lazily casting view of the same bf16 device tensors), so both
sides see identical parameters and no second 23.8 GB copy is made for bf16.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import flux as OF
from . import vae as OV


class CastingParams(dict):
    """Dict view of device parameters that hands out tensors in `dtype` (cast per access: fp32 needs no resident copy)."""

    def __init__(self, params: Dict[str, torch.Tensor], dtype: torch.dtype):
        super().__init__(params)
        self.dtype = dtype

    def __getitem__(self, k):
        v = super().__getitem__(k)
        return v if v.dtype == self.dtype else v.to(self.dtype)


def no_tf32():
    """fp32 means fp32: the oracle legs must not silently run TF32 tensor-core math."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@torch.no_grad()
def flux_forward(params, cfg: OF.FluxConfig, x, ctx, pooled, t, g, h2: int, w2: int, dtype=torch.bfloat16):
    """One MMDiT forward of the oracle on the device of `x`: x [B,S_img,C_in], ctx [B,S_txt,txt_dim], pooled [B,P],
    t / g fp32 [B] -> v [B,S_img,64] in `dtype`."""
    if dtype == torch.float32:
        no_tf32()
    p = CastingParams(params, dtype)
    img_ids, txt_ids = OF.image_ids(h2, w2), torch.zeros(ctx.shape[1], 3)
    return OF.flux_forward(p, cfg, x.to(dtype), ctx.to(dtype), pooled.to(dtype), t.float(), g.float(), img_ids, txt_ids)


@torch.no_grad()
def sample(params, cfg: OF.FluxConfig, latents_packed, ctx, pooled, guidance: float, num_steps: int, h2: int, w2: int,
           extra_cond: Optional[torch.Tensor] = None, start_step: int = 0, dtype=torch.bfloat16):
    """The oracle's flow-match Euler loop on the GPU (latents kept in `dtype` between steps, update in fp32 like the
    scheduler)."""
    if dtype == torch.float32:
        no_tf32()
    p = CastingParams(params, dtype)
    return OF.sample(p, cfg, latents_packed.to(dtype), ctx.to(dtype), pooled.to(dtype), guidance, num_steps, h2, w2,
                     extra_cond=None if extra_cond is None else extra_cond.to(dtype), start_step=start_step)


@torch.no_grad()
def vae_decode_u8(latents, p_vae, dtype=torch.float32):
    """Pipeline tail: latents [B,16,h,w] -> uint8 [B,H,W,3] with the oracle VAE on the device of `latents`."""
    if dtype == torch.float32:
        no_tf32()
    p = CastingParams({k: v.to(latents.device) for k, v in p_vae.items()}, dtype)
    return OV.postprocess_u8(OV.decode_latents(latents.to(dtype), p).float())


@torch.no_grad()
def vae_encode(x_u8_nhwc, p_vae, noise=None, mask=None, dtype=torch.float32):
    """uint8 [B,H,W,3] (+ optional repaint mask [B,H,W] in {0,1}: masked pixels -> 0 after normalisation, like
    FluxFillPipeline's masked_image) -> scaled latents [B,16,H/8,W/8]."""
    if dtype == torch.float32:
        no_tf32()
    p = CastingParams({k: v.to(x_u8_nhwc.device) for k, v in p_vae.items()}, dtype)
    img = OV.preprocess_image(x_u8_nhwc).to(dtype)
    if mask is not None:
        img = img * (1.0 - mask[:, None].to(dtype))
    return OV.encode_image(img, p, None if noise is None else noise.to(dtype))
