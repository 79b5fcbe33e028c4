"""Oracle: Flux VAE (AutoencoderKL of FLUX.1-dev / Fill-dev) encode / decode and the pipelines' image
pre/post-processing (PyTorch fp32, CPU). TEST INFRASTRUCTURE ONLY.

What the reference reaches through `pipe(...).images` (batch_generate_flux_kshot.py:467-474) and
`pipe_fill(image=..., mask_image=...)` (outpainting_updown_sampling_redux.py:1246-1257): diffusers==0.33.1
`AutoencoderKL` (block_out_channels 128/256/512/512, 2 layers per block, GroupNorm-32 eps 1e-6, SiLU, one
single-head attention in the mid block, 16 latent channels, scaling 0.3611, shift 0.1159, no quant convs) and
`VaeImageProcessor`. diffusers is NOT in /root/reference and not installable offline: PARITY UNPINNED by the
reference. This restates the published architecture and is cross-checked in tests/test_vae_oracle.py against
the independent BFL-style implementation shipped in this image (torchtitan.experiments.flux.model.autoencoder).

Parameters: flat dict of fp32 tensors (conv weights [Cout,Cin,kh,kw]):
  {enc,dec}.conv_in / conv_out / norm_out, {enc,dec}.mid.res{0,1}.*, {enc,dec}.mid.attn.{norm,q,k,v,proj},
  enc.down{L}.res{i}.*, enc.down{L}.downsample, dec.up{L}.res{i}.*, dec.up{L}.upsample   (L in execution order),
  res block = norm1, conv1, norm2, conv2 (+ short when the channel count changes).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

CH = 128
CH_MULT = (1, 2, 4, 4)
Z_CHANNELS = 16
SCALE_FACTOR = 0.3611
SHIFT_FACTOR = 0.1159
GN_GROUPS, GN_EPS = 32, 1e-6


def _conv_shapes(prefix, cin, cout, k=3):
    return {prefix + ".w": (cout, cin, k, k), prefix + ".b": (cout,)}


def _res_shapes(prefix, cin, cout):
    s = {prefix + ".norm1.w": (cin,), prefix + ".norm1.b": (cin,), prefix + ".norm2.w": (cout,), prefix + ".norm2.b": (cout,)}
    s.update(_conv_shapes(prefix + ".conv1", cin, cout))
    s.update(_conv_shapes(prefix + ".conv2", cout, cout))
    if cin != cout:
        s.update(_conv_shapes(prefix + ".short", cin, cout, 1))
    return s


def _attn_shapes(prefix, c):
    s = {prefix + ".norm.w": (c,), prefix + ".norm.b": (c,)}
    for n in ("q", "k", "v", "proj"):
        s.update(_conv_shapes(f"{prefix}.{n}", c, c, 1))
    return s


def param_shapes(ch: int = CH, ch_mult=CH_MULT, z: int = Z_CHANNELS) -> Dict[str, tuple]:
    s = {}
    # encoder
    s.update(_conv_shapes("enc.conv_in", 3, ch))
    cin = ch
    for L, m in enumerate(ch_mult):
        cout = ch * m
        for i in range(2):
            s.update(_res_shapes(f"enc.down{L}.res{i}", cin, cout))
            cin = cout
        if L != len(ch_mult) - 1:
            s.update(_conv_shapes(f"enc.down{L}.downsample", cin, cin))
    for side, c in (("enc", cin), ("dec", ch * ch_mult[-1])):
        s.update(_res_shapes(f"{side}.mid.res0", c, c))
        s.update(_attn_shapes(f"{side}.mid.attn", c))
        s.update(_res_shapes(f"{side}.mid.res1", c, c))
    s.update({"enc.norm_out.w": (cin,), "enc.norm_out.b": (cin,)})
    s.update(_conv_shapes("enc.conv_out", cin, 2 * z))
    # decoder (levels in execution order: widest channels first)
    cin = ch * ch_mult[-1]
    s.update(_conv_shapes("dec.conv_in", z, cin))
    for L, m in enumerate(reversed(ch_mult)):
        cout = ch * m
        for i in range(3):
            s.update(_res_shapes(f"dec.up{L}.res{i}", cin, cout))
            cin = cout
        if L != len(ch_mult) - 1:
            s.update(_conv_shapes(f"dec.up{L}.upsample", cin, cin))
    s.update({"dec.norm_out.w": (cin,), "dec.norm_out.b": (cin,)})
    s.update(_conv_shapes("dec.conv_out", cin, 3))
    return s


def init_params(seed: int = 5000, ch: int = CH, ch_mult=CH_MULT, z: int = Z_CHANNELS) -> Dict[str, torch.Tensor]:
    """Seeded stand-in for the checkpoint: conv std = fan_in^-1/2 (second conv of a res block and attention
    projections 0.5x so the residual stream stays O(1)), biases 0.02, GroupNorm gamma ~ 1, beta small."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes(ch, ch_mult, z).items():
        if ".norm" in name:
            p[name] = 1 + 0.1 * torch.randn(shape, generator=g) if name.endswith(".w") else 0.05 * torch.randn(shape, generator=g)
        elif name.endswith(".w"):
            fan_in = shape[1] * shape[2] * shape[3]
            damp = 0.5 if (".conv2." in name or ".proj." in name) else 1.0
            p[name] = torch.randn(shape, generator=g) * fan_in ** -0.5 * damp
        else:
            p[name] = 0.02 * torch.randn(shape, generator=g)
    return p


def _gn_silu(x, p, prefix, silu=True):
    y = F.group_norm(x, GN_GROUPS, p[prefix + ".w"], p[prefix + ".b"], GN_EPS)
    return F.silu(y) if silu else y


def _conv(x, p, prefix, stride=1, padding=1):
    return F.conv2d(x, p[prefix + ".w"], p[prefix + ".b"], stride=stride, padding=padding)


def res_block(x, p, prefix):
    h = _conv(_gn_silu(x, p, prefix + ".norm1"), p, prefix + ".conv1")
    h = _conv(_gn_silu(h, p, prefix + ".norm2"), p, prefix + ".conv2")
    if prefix + ".short.w" in p:
        x = _conv(x, p, prefix + ".short", padding=0)
    return x + h


def attn_block(x, p, prefix):
    B, C, H, W = x.shape
    h = _gn_silu(x, p, prefix + ".norm", silu=False)
    q, k, v = (_conv(h, p, f"{prefix}.{n}", padding=0).reshape(B, C, H * W).transpose(1, 2) for n in ("q", "k", "v"))
    a = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1) @ v          # single head, head dim = C
    a = a.transpose(1, 2).reshape(B, C, H, W)
    return x + _conv(a, p, prefix + ".proj", padding=0)


def _levels(p, side):
    n = 0
    while f"{side}{n}.res0.norm1.w" in p:
        n += 1
    return n


def decoder(z, p):
    """z [B,16,h,w] (already un-scaled) -> image [B,3,8h,8w] in roughly [-1,1]."""
    h = _conv(z, p, "dec.conv_in")
    h = res_block(h, p, "dec.mid.res0")
    h = attn_block(h, p, "dec.mid.attn")
    h = res_block(h, p, "dec.mid.res1")
    n = _levels(p, "dec.up")
    for L in range(n):
        for i in range(3):
            h = res_block(h, p, f"dec.up{L}.res{i}")
        if L != n - 1:
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), p, f"dec.up{L}.upsample")
    return _conv(_gn_silu(h, p, "dec.norm_out"), p, "dec.conv_out")


def encoder(x, p):
    """x [B,3,H,W] in [-1,1] -> moments [B,32,H/8,W/8] = cat(mean, logvar)."""
    h = _conv(x, p, "enc.conv_in")
    n = _levels(p, "enc.down")
    for L in range(n):
        for i in range(2):
            h = res_block(h, p, f"enc.down{L}.res{i}")
        if L != n - 1:
            h = _conv(F.pad(h, (0, 1, 0, 1)), p, f"enc.down{L}.downsample", stride=2, padding=0)
    h = res_block(h, p, "enc.mid.res0")
    h = attn_block(h, p, "enc.mid.attn")
    h = res_block(h, p, "enc.mid.res1")
    return _conv(_gn_silu(h, p, "enc.norm_out"), p, "enc.conv_out")


def decode_latents(latents, p):
    """What the Flux pipelines do after the loop: z / scaling + shift -> decoder."""
    return decoder(latents / SCALE_FACTOR + SHIFT_FACTOR, p)


def encode_image(x, p, noise=None):
    """(sample(moments) - shift) * scaling; noise=None -> the distribution mode (mean)."""
    mean, logvar = encoder(x, p).chunk(2, dim=1)
    z = mean if noise is None else mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise
    return (z - SHIFT_FACTOR) * SCALE_FACTOR


def postprocess_u8(img):
    """VaeImageProcessor.postprocess(output_type="pil") up to the PIL wrap: denormalise, clamp, NHWC, x255 round."""
    x = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1)
    return (x * 255).round().to(torch.uint8)


def preprocess_image(u8_nhwc):
    """uint8 [B,H,W,3] -> float [B,3,H,W] in [-1,1] (VaeImageProcessor.preprocess without resizing)."""
    return u8_nhwc.permute(0, 3, 1, 2).float() / 255.0 * 2.0 - 1.0
