"""SigLIP vision tower + Redux image embedder on the sm_100a kernels - the image-prompt half of
FluxPriorReduxPipeline (reference call sites batch_generate_flux_kshot.py:459-465,
outpainting_updown_sampling_redux.py:1237-1243): 384^2 image -> 729 SigLIP tokens (hidden 1152) ->
ReduxImageEncoder (Linear 1152->12288, SiLU, Linear 12288->4096) -> 729 image-prompt tokens.

Same kernels as the CLIP tower (tcgen05 GEMM with fused bias / GELU / SiLU / residual epilogues, tcgen05
attention, LayerNorm kernel). Two shape adaptations done once at weight load, in the layout only:
  * head dim 72 is zero-padded to 128 (the attention kernel's tile) and sqrt(128/72) is folded into the Q
    projection so the kernel's 1/sqrt(128) becomes 1/sqrt(72);
  * the MLP width 4304 is zero-padded to 4320 (GEMM N granularity 32).
State dict keys are Hugging Face SiglipVisionModel names. No PyTorch arithmetic on the path, no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib, ops

HD_PAD = 128


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
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def tokens(self) -> int:
        return self.grid ** 2


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class SiglipVisionTower:
    """SiglipVisionModel stand-in: `last_hidden_state(pixel_values)`."""

    def __init__(self, cfg: SiglipConfig, state: Dict[str, torch.Tensor], device="cuda"):
        _lib.load()
        self.cfg, self.device = cfg, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SiglipVisionTower runs only on CUDA (sm_100a); there is no CPU path")
        s = {k: v.detach().float().cpu() for k, v in state.items()}
        bf = lambda t: t.to(self.device, torch.bfloat16).contiguous()  # noqa: E731
        w, H, p = cfg.hidden, cfg.heads, cfg.patch
        hd = w // H
        assert hd <= HD_PAD and w % 8 == 0
        k = 3 * p * p
        self.kpad = _pad(k, 8)
        conv = torch.zeros(w, self.kpad)
        conv[:, :k] = s["vision_model.embeddings.patch_embedding.weight"].reshape(w, k)
        self.conv_w, self.conv_b = bf(conv), bf(s["vision_model.embeddings.patch_embedding.bias"])
        self.pos = bf(s["vision_model.embeddings.position_embedding.weight"])
        self.ln_post = (bf(s["vision_model.post_layernorm.weight"]), bf(s["vision_model.post_layernorm.bias"]))
        self.mlp_pad = _pad(cfg.mlp, 32)
        qscale = (HD_PAD / hd) ** 0.5
        self.blocks = []
        for i in range(cfg.layers):
            q = f"vision_model.encoder.layers.{i}."
            wqkv = torch.zeros(3, H, HD_PAD, w)
            bqkv = torch.zeros(3, H, HD_PAD)
            for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
                sc = qscale if j == 0 else 1.0
                wqkv[j, :, :hd] = s[q + f"self_attn.{n}.weight"].view(H, hd, w) * sc
                bqkv[j, :, :hd] = s[q + f"self_attn.{n}.bias"].view(H, hd) * sc
            wo = torch.zeros(w, H, HD_PAD)
            wo[:, :, :hd] = s[q + "self_attn.out_proj.weight"].view(w, H, hd)
            fc1 = torch.zeros(self.mlp_pad, w)
            fc1[:cfg.mlp] = s[q + "mlp.fc1.weight"]
            b1 = torch.zeros(self.mlp_pad)
            b1[:cfg.mlp] = s[q + "mlp.fc1.bias"]
            fc2 = torch.zeros(w, self.mlp_pad)
            fc2[:, :cfg.mlp] = s[q + "mlp.fc2.weight"]
            self.blocks.append({
                "ln1": (bf(s[q + "layer_norm1.weight"]), bf(s[q + "layer_norm1.bias"])),
                "ln2": (bf(s[q + "layer_norm2.weight"]), bf(s[q + "layer_norm2.bias"])),
                "wqkv": bf(wqkv.reshape(3 * H * HD_PAD, w)), "bqkv": bf(bqkv.reshape(-1)),
                "wo": bf(wo.reshape(w, H * HD_PAD)), "bo": bf(s[q + "self_attn.out_proj.bias"]),
                "fc1": bf(fc1), "b1": bf(b1), "fc2": bf(fc2), "b2": bf(s[q + "mlp.fc2.bias"])})
        self._pos_tiled = {}

    @torch.no_grad()
    def last_hidden_state(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """pixel_values float [B,3,R,R] in [-1,1] on the GPU -> bf16 [B, tokens, hidden]."""
        cfg, lib = self.cfg, _lib.load()
        if not pixel_values.is_cuda:
            raise RuntimeError("last_hidden_state: input must be a CUDA tensor (no CPU path)")
        x = pixel_values.to(torch.float32).contiguous()
        B, w, H, L, g = x.shape[0], cfg.hidden, cfg.heads, cfg.tokens, cfg.grid
        if x.shape[1:] != (3, cfg.image, cfg.image):
            raise ValueError(f"expected [B,3,{cfg.image},{cfg.image}], got {tuple(x.shape)}")
        dev, st = x.device, _lib.current_stream_ptr(x.device)
        patches = torch.empty((B * L, self.kpad), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.drag_vit_patchify(_lib.ptr(x), _lib.ptr(patches), B, cfg.image, cfg.patch, self.kpad, st),
                   "drag_vit_patchify")
        if B not in self._pos_tiled:
            self._pos_tiled = {B: self.pos.repeat(B, 1).contiguous()}
        h = ops.linear(patches, self.conv_w, self.conv_b, mode=ops.EPI_GATE_RESID, resid=self._pos_tiled[B])
        y = torch.empty_like(h)
        q = torch.empty((B, H, L, HD_PAD), dtype=torch.bfloat16, device=dev)
        k, v = torch.empty_like(q), torch.empty_like(q)
        a = torch.empty((B * L, H * HD_PAD), dtype=torch.bfloat16, device=dev)
        u = torch.empty((B * L, self.mlp_pad), dtype=torch.bfloat16, device=dev)
        for blk in self.blocks:
            ops.layernorm(h, blk["ln1"][0], blk["ln1"][1], eps=cfg.eps, out=y)
            _lib.check(lib.drag_gemm_qkv_split(_lib.ptr(y), w, _lib.ptr(blk["wqkv"]), w, B * L, w, H, HD_PAD,
                                               _lib.ptr(blk["bqkv"]), _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), L, 0, L, st),
                       "drag_gemm_qkv_split")
            ops.attention(q, k, v, 0, out1=a)
            ops.linear(a, blk["wo"], blk["bo"], mode=ops.EPI_GATE_RESID, resid=h, out=h)
            ops.layernorm(h, blk["ln2"][0], blk["ln2"][1], eps=cfg.eps, out=y)
            ops.linear(y, blk["fc1"], blk["b1"], mode=ops.EPI_GELU_TANH, out=u)
            ops.linear(u, blk["fc2"], blk["b2"], mode=ops.EPI_GATE_RESID, resid=h, out=h)
        return ops.layernorm(h, self.ln_post[0], self.ln_post[1], eps=cfg.eps).view(B, L, w)


class ReduxImageEncoder:
    """redux_down(silu(redux_up(tokens))) - two GEMMs, SiLU in the first epilogue."""

    def __init__(self, state: Dict[str, torch.Tensor], device="cuda"):
        dev = torch.device(device)
        bf = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()  # noqa: E731
        self.up_w, self.up_b = bf(state["redux_up.weight"]), bf(state["redux_up.bias"])
        self.down_w, self.down_b = bf(state["redux_down.weight"]), bf(state["redux_down.bias"])

    @torch.no_grad()
    def __call__(self, tokens: torch.Tensor) -> torch.Tensor:
        B, L, w = tokens.shape
        hid = ops.linear(tokens.reshape(B * L, w).contiguous(), self.up_w, self.up_b, mode=ops.EPI_SILU)
        return ops.linear(hid, self.down_w, self.down_b).view(B, L, -1)


def preprocess(pil_images, size: int = 384) -> torch.Tensor:
    """SiglipImageProcessor on the host: RGB, bicubic resize to size x size, /255, (x - 0.5) / 0.5 -> [B,3,size,size]."""
    import numpy as np
    from PIL import Image
    out = []
    for im in pil_images:
        a = np.asarray(im.convert("RGB").resize((size, size), Image.BICUBIC), dtype=np.float32) / 255.0
        out.append(torch.from_numpy((a - 0.5) / 0.5).permute(2, 0, 1))
    return torch.stack(out)
