"""CLIP ViT image encoder on the sm_100a kernels - drop-in for the reference's OpenAI-clip calls:

    model, preprocess = clip.load("ViT-B/32", device=device)             # retrieval/...:209
    image_embedding = model.encode_image(preprocess(image).unsqueeze(0).to(device))   # :168-171
    image_embedding = image_embedding / image_embedding.norm(dim=-1, keepdim=True)   # :172

`load` returns (model, preprocess) with the same call shapes. Weights: an OpenAI-format state dict
(`visual.*` keys, e.g. from a real checkpoint) or, with no checkpoints offline, a seeded random init.
Every matmul runs on the tcgen05 GEMM (patch embedding as patchify + GEMM since stride == kernel),
attention on the tcgen05 attention kernel (head dim 64), LayerNorm / QuickGELU / residual adds are
fused kernels or GEMM epilogues. No PyTorch arithmetic and no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib, ops

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclass
class ViTConfig:
    width: int = 768
    layers: int = 12
    heads: int = 12
    patch: int = 32
    image: int = 224
    out_dim: int = 512

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def tokens(self) -> int:
        return self.grid ** 2 + 1


CONFIGS = {"ViT-B/32": ViTConfig(768, 12, 12, 32, 224, 512), "ViT-B/16": ViTConfig(768, 12, 12, 16, 224, 512),
           "ViT-L/14": ViTConfig(1024, 24, 16, 14, 224, 768)}


def available_models():
    return list(CONFIGS)


def random_state(cfg: ViTConfig, seed: int = 2000) -> Dict[str, torch.Tensor]:
    """Seeded stand-in for the OpenAI checkpoint (same key layout)."""
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


class CLIPVisual:
    """Image tower of CLIP; `encode_image` mirrors clip.model.CLIP.encode_image."""

    def __init__(self, cfg: ViTConfig, state: Dict[str, torch.Tensor], device):
        _lib.load()
        self.cfg, self.device = cfg, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CLIPVisual runs only on CUDA (sm_100a); there is no CPU path")
        bf = lambda t: t.to(self.device, torch.bfloat16).contiguous()  # noqa: E731
        w, p = cfg.width, cfg.patch
        k = 3 * p * p
        self.kpad = (k + 7) // 8 * 8
        conv = torch.zeros(w, self.kpad)
        conv[:, :k] = state["visual.conv1.weight"].float().reshape(w, k)      # columns (c, py, px)
        self.conv_w = bf(conv)
        self.cls = bf(state["visual.class_embedding"])
        self.pos = bf(state["visual.positional_embedding"])
        self.ln_pre = (bf(state["visual.ln_pre.weight"]), bf(state["visual.ln_pre.bias"]))
        self.ln_post = (bf(state["visual.ln_post.weight"]), bf(state["visual.ln_post.bias"]))
        self.proj_t = bf(state["visual.proj"].float().t())                      # [out, w]
        self.blocks = []
        for i in range(cfg.layers):
            q = f"visual.transformer.resblocks.{i}."
            self.blocks.append({n: bf(state[q + n]) for n in (
                "ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_weight", "attn.in_proj_bias",
                "attn.out_proj.weight", "attn.out_proj.bias", "mlp.c_fc.weight", "mlp.c_fc.bias",
                "mlp.c_proj.weight", "mlp.c_proj.bias")})

    def eval(self):
        return self

    @torch.no_grad()
    def encode_image(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """image [B,3,R,R] float (already CLIP-normalised, as `preprocess` produces) -> float32 [B,out_dim]."""
        cfg, lib = self.cfg, _lib.load()
        if not image.is_cuda:
            raise RuntimeError("encode_image: input must be a CUDA tensor (no CPU path)")
        x = image.to(torch.float32).contiguous()
        B, w, H, L, g = x.shape[0], cfg.width, cfg.heads, cfg.tokens, cfg.grid
        if x.shape[1:] != (3, cfg.image, cfg.image):
            raise ValueError(f"expected [B,3,{cfg.image},{cfg.image}], got {tuple(x.shape)}")
        dev, st = x.device, _lib.current_stream_ptr(x.device)
        patches = torch.empty((B * g * g, self.kpad), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.drag_vit_patchify(_lib.ptr(x), _lib.ptr(patches), B, cfg.image, cfg.patch, self.kpad, st),
                   "drag_vit_patchify")
        pe = ops.linear(patches, self.conv_w)
        h = torch.empty((B * L, w), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.drag_vit_assemble(_lib.ptr(pe), _lib.ptr(self.cls), _lib.ptr(self.pos), _lib.ptr(h), B, g * g,
                                         w, st), "drag_vit_assemble")
        ops.layernorm(h, self.ln_pre[0], self.ln_pre[1], eps=1e-5, out=h)
        y = torch.empty_like(h)
        q = torch.empty((B, H, L, w // H), dtype=torch.bfloat16, device=dev)
        k, v = torch.empty_like(q), torch.empty_like(q)
        a = torch.empty_like(h)
        u = torch.empty((B * L, 4 * w), dtype=torch.bfloat16, device=dev)
        for blk in self.blocks:
            ops.layernorm(h, blk["ln_1.weight"], blk["ln_1.bias"], eps=1e-5, out=y)
            _lib.check(lib.drag_gemm_qkv_split(_lib.ptr(y), w, _lib.ptr(blk["attn.in_proj_weight"]), w, B * L, w, H,
                                               w // H, _lib.ptr(blk["attn.in_proj_bias"]), _lib.ptr(q), _lib.ptr(k),
                                               _lib.ptr(v), L, 0, L, st), "drag_gemm_qkv_split")
            ops.attention(q, k, v, 0, out1=a)
            ops.linear(a, blk["attn.out_proj.weight"], blk["attn.out_proj.bias"], mode=ops.EPI_GATE_RESID, resid=h, out=h)
            ops.layernorm(h, blk["ln_2.weight"], blk["ln_2.bias"], eps=1e-5, out=y)
            ops.linear(y, blk["mlp.c_fc.weight"], blk["mlp.c_fc.bias"], mode=ops.EPI_QUICK_GELU, out=u)
            ops.linear(u, blk["mlp.c_proj.weight"], blk["mlp.c_proj.bias"], mode=ops.EPI_GATE_RESID, resid=h, out=h)
        cls_rows = h.view(B, L, w)[:, 0, :]                                   # strided view, ld = L*w
        c = ops.layernorm(cls_rows, self.ln_post[0], self.ln_post[1], eps=1e-5)
        emb = ops.linear(c, self.proj_t, None, mode=ops.EPI_BIAS_F32)
        return ops.l2_normalize(emb) if normalize else emb


class CLIP:
    """What clip.load returns: only the image tower is on the reference's hot path."""

    def __init__(self, visual: CLIPVisual):
        self.visual = visual

    def eval(self):
        return self

    def encode_image(self, image, normalize: bool = False):
        return self.visual.encode_image(image, normalize=normalize)


def _transform(n_px: int):
    """clip._transform: Resize(bicubic, short side) -> CenterCrop -> RGB -> ToTensor -> Normalize."""
    from torchvision import transforms as T
    return T.Compose([T.Resize(n_px, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(n_px),
                      lambda im: im.convert("RGB"), T.ToTensor(), T.Normalize(CLIP_MEAN, CLIP_STD)])


def load(name: str = "ViT-B/32", device="cuda", state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 2000):
    """(model, preprocess) like clip.load(name, device). `state_dict` takes OpenAI `visual.*` keys."""
    if name not in CONFIGS:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    cfg = CONFIGS[name]
    state = state_dict if state_dict is not None else random_state(cfg, seed)
    return CLIP(CLIPVisual(cfg, state, device)), _transform(cfg.image)
