"""Seeded random-init parameter sets for the models that have no checkpoint offline (VAE, SigLIP, Redux).
The draws mirror oracle/vae.py::init_params and oracle/siglip.py::init_state / init_redux call for call, so the
GPU parity tests can build the same weights on both sides from a seed (tests/test_models_init.py checks equality)."""
from __future__ import annotations

from typing import Dict

import torch

CH, CH_MULT, Z_CHANNELS = 128, (1, 2, 4, 4), 16


def _conv(s, prefix, cin, cout, k=3):
    s[prefix + ".w"] = (cout, cin, k, k)
    s[prefix + ".b"] = (cout,)


def _res(s, prefix, cin, cout):
    s[prefix + ".norm1.w"], s[prefix + ".norm1.b"] = (cin,), (cin,)
    s[prefix + ".norm2.w"], s[prefix + ".norm2.b"] = (cout,), (cout,)
    _conv(s, prefix + ".conv1", cin, cout)
    _conv(s, prefix + ".conv2", cout, cout)
    if cin != cout:
        _conv(s, prefix + ".short", cin, cout, 1)


def _attn(s, prefix, c):
    s[prefix + ".norm.w"], s[prefix + ".norm.b"] = (c,), (c,)
    for n in ("q", "k", "v", "proj"):
        _conv(s, f"{prefix}.{n}", c, c, 1)


def param_shapes(ch: int = CH, ch_mult=CH_MULT, z: int = Z_CHANNELS) -> Dict[str, tuple]:
    s: Dict[str, tuple] = {}
    _conv(s, "enc.conv_in", 3, ch)
    cin = ch
    for L, m in enumerate(ch_mult):
        for i in range(2):
            _res(s, f"enc.down{L}.res{i}", cin, ch * m)
            cin = ch * m
        if L != len(ch_mult) - 1:
            _conv(s, f"enc.down{L}.downsample", cin, cin)
    for side, c in (("enc", cin), ("dec", ch * ch_mult[-1])):
        _res(s, f"{side}.mid.res0", c, c)
        _attn(s, f"{side}.mid.attn", c)
        _res(s, f"{side}.mid.res1", c, c)
    s["enc.norm_out.w"], s["enc.norm_out.b"] = (cin,), (cin,)
    _conv(s, "enc.conv_out", cin, 2 * z)
    cin = ch * ch_mult[-1]
    _conv(s, "dec.conv_in", z, cin)
    for L, m in enumerate(reversed(ch_mult)):
        for i in range(3):
            _res(s, f"dec.up{L}.res{i}", cin, ch * m)
            cin = ch * m
        if L != len(ch_mult) - 1:
            _conv(s, f"dec.up{L}.upsample", cin, cin)
    s["dec.norm_out.w"], s["dec.norm_out.b"] = (cin,), (cin,)
    _conv(s, "dec.conv_out", cin, 3)
    return s


def init_params(seed: int = 5000, ch: int = CH, ch_mult=CH_MULT, z: int = Z_CHANNELS) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes(ch, ch_mult, z).items():
        if ".norm" in name:
            p[name] = 1 + 0.1 * torch.randn(shape, generator=g) if name.endswith(".w") else 0.05 * torch.randn(shape, generator=g)
        elif name.endswith(".w"):
            damp = 0.5 if (".conv2." in name or ".proj." in name) else 1.0
            p[name] = torch.randn(shape, generator=g) * (shape[1] * shape[2] * shape[3]) ** -0.5 * damp
        else:
            p[name] = 0.02 * torch.randn(shape, generator=g)
    return p


def init_siglip(cfg, seed: int = 6000) -> Dict[str, torch.Tensor]:
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


def init_redux(seed: int = 6100, d_in: int = 1152, d_hidden: int = 3 * 4096, d_out: int = 4096):
    g = torch.Generator().manual_seed(seed)
    return {"redux_up.weight": torch.randn(d_hidden, d_in, generator=g) * d_in ** -0.5,
            "redux_up.bias": 0.02 * torch.randn(d_hidden, generator=g),
            "redux_down.weight": torch.randn(d_out, d_hidden, generator=g) * d_hidden ** -0.5,
            "redux_down.bias": 0.02 * torch.randn(d_out, generator=g)}
