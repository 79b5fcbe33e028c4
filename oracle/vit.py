"""Oracle: OpenAI CLIP VisionTransformer image tower (PyTorch fp32, CPU). TEST INFRASTRUCTURE ONLY.

The reference calls clip.load("ViT-B/32") and model.encode_image(x) followed by x / x.norm(dim=-1)
(retrieval/clip100_resnet_style_all_shots.py:209, :171-172, :284-285). The `clip` package
(openai/CLIP @ dcba3cb2, requirements.txt:4) is NOT in /root/reference and not installable offline:
PARITY UNPINNED by the reference. This restates clip/model.py::VisionTransformer (conv1 without bias,
class token, learned positions, ln_pre, pre-LN residual blocks with nn.MultiheadAttention and QuickGELU
MLP, ln_post on the class token, projection) on an OpenAI-format state dict, and is cross-checked in
tests/test_vit_oracle.py against transformers.CLIPVisionModelWithProjection (independent code, same math).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F


@dataclass
class ViTConfig:
    width: int = 768
    layers: int = 12
    heads: int = 12
    patch: int = 32
    image: int = 224
    out_dim: int = 512

    @property
    def tokens(self) -> int:
        return (self.image // self.patch) ** 2 + 1


CONFIGS = {"ViT-B/32": ViTConfig(768, 12, 12, 32, 224, 512), "ViT-B/16": ViTConfig(768, 12, 12, 16, 224, 512),
           "ViT-L/14": ViTConfig(1024, 24, 16, 14, 224, 768)}


def init_state(cfg: ViTConfig, seed: int = 2000) -> Dict[str, torch.Tensor]:
    """Random OpenAI-format state dict (visual.* keys): std = fan_in^-1/2 for matrices, LN gamma 1 beta 0,
    embeddings scaled width^-1/2 as in clip/model.py."""
    g = torch.Generator().manual_seed(seed)
    w, sc = cfg.width, cfg.width ** -0.5
    s = {"visual.conv1.weight": torch.randn(w, 3, cfg.patch, cfg.patch, generator=g) * (3 * cfg.patch ** 2) ** -0.5,
         "visual.class_embedding": sc * torch.randn(w, generator=g),
         "visual.positional_embedding": sc * torch.randn(cfg.tokens, w, generator=g),
         "visual.ln_pre.weight": torch.ones(w), "visual.ln_pre.bias": torch.zeros(w),
         "visual.ln_post.weight": torch.ones(w), "visual.ln_post.bias": torch.zeros(w),
         "visual.proj": sc * torch.randn(w, cfg.out_dim, generator=g)}
    for i in range(cfg.layers):
        p = f"visual.transformer.resblocks.{i}."
        s.update({p + "ln_1.weight": 1 + 0.05 * torch.randn(w, generator=g), p + "ln_1.bias": 0.02 * torch.randn(w, generator=g),
                  p + "ln_2.weight": 1 + 0.05 * torch.randn(w, generator=g), p + "ln_2.bias": 0.02 * torch.randn(w, generator=g),
                  p + "attn.in_proj_weight": torch.randn(3 * w, w, generator=g) * sc,
                  p + "attn.in_proj_bias": 0.02 * torch.randn(3 * w, generator=g),
                  p + "attn.out_proj.weight": torch.randn(w, w, generator=g) * sc,
                  p + "attn.out_proj.bias": 0.02 * torch.randn(w, generator=g),
                  p + "mlp.c_fc.weight": torch.randn(4 * w, w, generator=g) * sc,
                  p + "mlp.c_fc.bias": 0.02 * torch.randn(4 * w, generator=g),
                  p + "mlp.c_proj.weight": torch.randn(w, 4 * w, generator=g) * (4 * w) ** -0.5,
                  p + "mlp.c_proj.bias": 0.02 * torch.randn(w, generator=g)})
    return s


def encode_image(state: Dict[str, torch.Tensor], cfg: ViTConfig, x: torch.Tensor) -> torch.Tensor:
    """x [B,3,R,R] (already normalised) -> [B,out_dim], un-normalised, like model.encode_image."""
    s = {k: v.float() for k, v in state.items()}
    B, w, H = x.shape[0], cfg.width, cfg.heads
    h = F.conv2d(x.float(), s["visual.conv1.weight"], None, stride=cfg.patch)         # [B,w,g,g]
    h = h.reshape(B, w, -1).permute(0, 2, 1)                                             # [B,g^2,w]
    cls = s["visual.class_embedding"].expand(B, 1, w)
    h = torch.cat([cls, h], 1) + s["visual.positional_embedding"]
    h = F.layer_norm(h, (w,), s["visual.ln_pre.weight"], s["visual.ln_pre.bias"], 1e-5)
    L = h.shape[1]
    for i in range(cfg.layers):
        p = f"visual.transformer.resblocks.{i}."
        y = F.layer_norm(h, (w,), s[p + "ln_1.weight"], s[p + "ln_1.bias"], 1e-5)
        qkv = F.linear(y, s[p + "attn.in_proj_weight"], s[p + "attn.in_proj_bias"]).view(B, L, 3, H, w // H)
        q, k, v = (qkv[:, :, j].permute(0, 2, 1, 3) for j in range(3))
        a = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B, L, w)
        h = h + F.linear(a, s[p + "attn.out_proj.weight"], s[p + "attn.out_proj.bias"])
        y = F.layer_norm(h, (w,), s[p + "ln_2.weight"], s[p + "ln_2.bias"], 1e-5)
        y = F.linear(y, s[p + "mlp.c_fc.weight"], s[p + "mlp.c_fc.bias"])
        y = y * torch.sigmoid(1.702 * y)                                                 # QuickGELU
        h = h + F.linear(y, s[p + "mlp.c_proj.weight"], s[p + "mlp.c_proj.bias"])
    c = F.layer_norm(h[:, 0], (w,), s["visual.ln_post.weight"], s["visual.ln_post.bias"], 1e-5)
    return c @ s["visual.proj"]


def embed(state, cfg, x):
    """encode_image + the reference's caller-side L2 normalisation (:172)."""
    e = encode_image(state, cfg, x)
    return e / e.norm(dim=-1, keepdim=True)
