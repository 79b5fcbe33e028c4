"""Oracle: SigLIP vision tower + Redux image embedder (PyTorch fp32, CPU). TEST INFRASTRUCTURE ONLY.

What FluxPriorReduxPipeline runs on each prompt image before the blend (reference call sites
batch_generate_flux_kshot.py:459-465, outpainting_updown_sampling_redux.py:1237-1243): diffusers==0.33.1 +
transformers SiglipVisionModel (google/siglip-so400m-patch14-384: hidden 1152, 27 layers, 16 heads of 72,
MLP 4304, patch 14, 384^2 -> 729 tokens, LayerNorm eps 1e-6, gelu_pytorch_tanh, no class token) followed by
ReduxImageEncoder (Linear 1152 -> 12288, SiLU, Linear 12288 -> 4096). Neither package is in /root/reference:
PARITY UNPINNED by the reference. Restated from the published architecture; cross-checked in
tests/test_siglip_oracle.py against transformers.SiglipVisionModel (independent code, same math).
State dict uses the Hugging Face key names, so a real checkpoint loads unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F


@dataclass
class SiglipConfig:
    hidden: int = 1152
    layers: int = 27
    heads: int = 16
    mlp: int = 4304
    patch: int = 14
    image: int = 384
    eps: float = 1e-6

    @property
    def tokens(self) -> int:
        return (self.image // self.patch) ** 2


SO400M = SiglipConfig()
REDUX_IN, REDUX_HIDDEN, REDUX_OUT = 1152, 4096 * 3, 4096


def init_state(cfg: SiglipConfig, seed: int = 6000) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    w, sc = cfg.hidden, cfg.hidden ** -0.5
    e = "vision_model.embeddings."
    s = {e + "patch_embedding.weight": torch.randn(w, 3, cfg.patch, cfg.patch, generator=g) * (3 * cfg.patch ** 2) ** -0.5,
         e + "patch_embedding.bias": 0.02 * torch.randn(w, generator=g),
         e + "position_embedding.weight": 0.3 * torch.randn(cfg.tokens, w, generator=g),
         "vision_model.post_layernorm.weight": 1 + 0.05 * torch.randn(w, generator=g),
         "vision_model.post_layernorm.bias": 0.02 * torch.randn(w, generator=g)}
    for i in range(cfg.layers):
        p = f"vision_model.encoder.layers.{i}."
        for n in ("layer_norm1", "layer_norm2"):
            s[p + n + ".weight"] = 1 + 0.05 * torch.randn(w, generator=g)
            s[p + n + ".bias"] = 0.02 * torch.randn(w, generator=g)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"] = torch.randn(w, w, generator=g) * sc
            s[p + f"self_attn.{n}.bias"] = 0.02 * torch.randn(w, generator=g)
        s[p + "mlp.fc1.weight"] = torch.randn(cfg.mlp, w, generator=g) * sc
        s[p + "mlp.fc1.bias"] = 0.02 * torch.randn(cfg.mlp, generator=g)
        s[p + "mlp.fc2.weight"] = torch.randn(w, cfg.mlp, generator=g) * cfg.mlp ** -0.5
        s[p + "mlp.fc2.bias"] = 0.02 * torch.randn(w, generator=g)
    return s


def last_hidden_state(state, cfg: SiglipConfig, x: torch.Tensor) -> torch.Tensor:
    """x [B,3,R,R] (normalised to [-1,1]) -> [B, tokens, hidden] = SiglipVisionModel(...).last_hidden_state."""
    s = {k: v.float() for k, v in state.items()}
    B, w, H = x.shape[0], cfg.hidden, cfg.heads
    e = "vision_model.embeddings."
    h = F.conv2d(x.float(), s[e + "patch_embedding.weight"], s[e + "patch_embedding.bias"], stride=cfg.patch)
    h = h.flatten(2).transpose(1, 2) + s[e + "position_embedding.weight"]
    L = h.shape[1]
    for i in range(cfg.layers):
        p = f"vision_model.encoder.layers.{i}."
        y = F.layer_norm(h, (w,), s[p + "layer_norm1.weight"], s[p + "layer_norm1.bias"], cfg.eps)
        q, k, v = (F.linear(y, s[p + f"self_attn.{n}.weight"], s[p + f"self_attn.{n}.bias"]).view(B, L, H, w // H).transpose(1, 2)
                   for n in ("q_proj", "k_proj", "v_proj"))
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, w)
        h = h + F.linear(a, s[p + "self_attn.out_proj.weight"], s[p + "self_attn.out_proj.bias"])
        y = F.layer_norm(h, (w,), s[p + "layer_norm2.weight"], s[p + "layer_norm2.bias"], cfg.eps)
        y = F.gelu(F.linear(y, s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"]), approximate="tanh")
        h = h + F.linear(y, s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"])
    return F.layer_norm(h, (w,), s["vision_model.post_layernorm.weight"], s["vision_model.post_layernorm.bias"], cfg.eps)


def init_redux(seed: int = 6100, d_in: int = REDUX_IN, d_hidden: int = REDUX_HIDDEN, d_out: int = REDUX_OUT):
    g = torch.Generator().manual_seed(seed)
    return {"redux_up.weight": torch.randn(d_hidden, d_in, generator=g) * d_in ** -0.5,
            "redux_up.bias": 0.02 * torch.randn(d_hidden, generator=g),
            "redux_down.weight": torch.randn(d_out, d_hidden, generator=g) * d_hidden ** -0.5,
            "redux_down.bias": 0.02 * torch.randn(d_out, generator=g)}


def redux_embed(redux, tokens: torch.Tensor) -> torch.Tensor:
    """ReduxImageEncoder: redux_down(silu(redux_up(x)))."""
    r = {k: v.float() for k, v in redux.items()}
    return F.linear(F.silu(F.linear(tokens.float(), r["redux_up.weight"], r["redux_up.bias"])), r["redux_down.weight"],
                    r["redux_down.bias"])


def preprocess(pil_images, size: int = 384) -> torch.Tensor:
    """SiglipImageProcessor: RGB, resize to size x size (bicubic), /255, (x - 0.5) / 0.5."""
    import numpy as np
    from PIL import Image
    out = []
    for im in pil_images:
        a = np.asarray(im.convert("RGB").resize((size, size), Image.BICUBIC), dtype=np.float32) / 255.0
        out.append(torch.from_numpy((a - 0.5) / 0.5).permute(2, 0, 1))
    return torch.stack(out)
