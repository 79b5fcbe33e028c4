"""CLIP ViT image encoder on the sm_100a kernels - drop-in for the reference's OpenAI-clip calls:

    model, preprocess = clip.load("ViT-B/32", device=device)             # retrieval/...:209
    image_embedding = model.encode_image(preprocess(image).unsqueeze(0).to(device))   # :168-171
    image_embedding = image_embedding / image_embedding.norm(dim=-1, keepdim=True)   # :172

`load` returns (model, preprocess) with the same call shapes. Weights: an OpenAI-format state dict
(`visual.*` keys, e.g. from a real checkpoint) or, with no checkpoints offline, a seeded random init.
The whole tower runs inside ONE C call per batch (drag_vit_encode, csrc/vit_engine.cu): every matmul on the
tcgen05 GEMM (patch embedding as patchify + GEMM since stride == kernel), attention on the tcgen05 attention
kernel (head dim 64), LayerNorm / QuickGELU / residual adds as fused kernels or GEMM epilogues. `encode_image`
takes the normalised float tensor `preprocess` returns (the reference contract) or raw uint8 pixels
(`preprocess_u8`: same resize / crop on the host, ToTensor + Normalize on the GPU - a quarter of the PCIe bytes,
bit-identical patches). No PyTorch arithmetic and no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

import ctypes as C

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


# pointer order of drag_vit_set_weights: 8 globals, then 12 per block
ENGINE_GLOBALS = ("conv_w", "cls", "pos", "ln_pre.weight", "ln_pre.bias", "ln_post.weight", "ln_post.bias", "proj_t")
ENGINE_BLOCK = ("ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight",
                "attn.out_proj.bias", "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight",
                "mlp.c_proj.bias")
# ... followed by the LayerNorm-folded operands of the block (CLIPVisual._fold): gamma-scaled weights (bf16) and fp32 s / c
ENGINE_FOLDED = ("qkv_wf", "qkv_s", "qkv_c", "fc_wf", "fc_s", "fc_c")
ENGINE_ORDER = ENGINE_GLOBALS + tuple("blocks.<i>." + n for n in ENGINE_BLOCK + ENGINE_FOLDED)


class _VitConfigC(C.Structure):
    _fields_ = [("width", C.c_int), ("layers", C.c_int), ("heads", C.c_int), ("patch", C.c_int), ("image", C.c_int),
                ("out_dim", C.c_int), ("max_batch", C.c_int), ("mean", C.c_float * 3), ("std", C.c_float * 3)]


class CLIPVisual:
    """Image tower of CLIP; `encode_image` mirrors clip.model.CLIP.encode_image. Owns the bf16 device weights and one C++
    engine (activation workspace for `max_batch` images; larger batches are chunked inside the library)."""

    def __init__(self, cfg: ViTConfig, state: Dict[str, torch.Tensor], device, max_batch: int = 512):
        lib = _lib.load()
        self.cfg, self.device = cfg, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CLIPVisual runs only on CUDA (sm_100a); there is no CPU path")
        bf = lambda t: t.to(self.device, torch.bfloat16).contiguous()  # noqa: E731
        w, p = cfg.width, cfg.patch
        k = 3 * p * p
        self.kpad = (k + 7) // 8 * 8
        conv = torch.zeros(w, self.kpad)
        conv[:, :k] = state["visual.conv1.weight"].float().reshape(w, k)      # columns (c, py, px)
        self.weights = {"conv_w": bf(conv), "cls": bf(state["visual.class_embedding"]),
                        "pos": bf(state["visual.positional_embedding"]),
                        "ln_pre.weight": bf(state["visual.ln_pre.weight"]), "ln_pre.bias": bf(state["visual.ln_pre.bias"]),
                        "ln_post.weight": bf(state["visual.ln_post.weight"]), "ln_post.bias": bf(state["visual.ln_post.bias"]),
                        "proj_t": bf(state["visual.proj"].float().t())}                 # [out, w]
        for i in range(cfg.layers):
            q = f"visual.transformer.resblocks.{i}."
            for n in ENGINE_BLOCK:
                self.weights[f"blocks.{i}.{n}"] = bf(state[q + n])
            for tag, ln, lin in (("qkv", "ln_1", "attn.in_proj_"), ("fc", "ln_2", "mlp.c_fc.")):
                wf, sv, cv = self._fold(state[q + ln + ".weight"], state[q + ln + ".bias"],
                                        state[q + lin + "weight"], state[q + lin + "bias"])
                self.weights[f"blocks.{i}.{tag}_wf"] = wf.to(self.device).contiguous()
                self.weights[f"blocks.{i}.{tag}_s"] = sv.to(self.device).contiguous()
                self.weights[f"blocks.{i}.{tag}_c"] = cv.to(self.device).contiguous()
        order = list(ENGINE_GLOBALS) + [f"blocks.{i}.{n}" for i in range(cfg.layers) for n in ENGINE_BLOCK + ENGINE_FOLDED]
        self.max_batch = int(max_batch)
        c = _VitConfigC(cfg.width, cfg.layers, cfg.heads, cfg.patch, cfg.image, cfg.out_dim, self.max_batch,
                        (C.c_float * 3)(*CLIP_MEAN), (C.c_float * 3)(*CLIP_STD))
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.drag_vit_create(C.byref(c), C.byref(self._h)), "drag_vit_create")
        ptrs = (C.c_void_p * len(order))(*[self.weights[n].data_ptr() for n in order])
        _lib.check(lib.drag_vit_set_weights(self._h, ptrs, len(order)), "drag_vit_set_weights")

    @staticmethod
    def _fold(gamma, beta, weight, bias):
        """LayerNorm(x; gamma, beta) @ W^T + b  ==  rstd * (x @ W'^T - mean * s) + c  with W' = W * gamma (rounded to bf16,
        the GEMM operand), s[n] = sum_k W'[n,k] (of the ROUNDED operand, so a constant row cancels exactly) and
        c[n] = sum_k beta[k] W[n,k] + b[n]. Both sides start from the bf16 weights the unfused path uses."""
        W = weight.to(torch.bfloat16).double()
        g, bt = gamma.to(torch.bfloat16).double(), beta.to(torch.bfloat16).double()
        wf = (W * g[None, :]).to(torch.bfloat16)
        s = wf.double().sum(1).float()
        c = (W @ bt + bias.to(torch.bfloat16).double()).float()
        return wf, s, c

    def fold_layernorm(self, on: bool = True) -> None:
        """A/B switch: True = ln_1 / ln_2 folded into the GEMMs around them, False (default) = separate LayerNorm kernels."""
        _lib.check(_lib.load().drag_vit_set_option(self._h, 1, int(on)), "drag_vit_set_option")

    def eval(self):
        return self

    @torch.no_grad()
    def encode_image(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """image [B,3,R,R]: float (already CLIP-normalised, as `preprocess` produces) or uint8 raw pixels (as
        `preprocess_u8` produces; normalised on the GPU) -> float32 [B,out_dim]."""
        cfg = self.cfg
        if not image.is_cuda:
            raise RuntimeError("encode_image: input must be a CUDA tensor (no CPU path)")
        if tuple(image.shape[1:]) != (3, cfg.image, cfg.image):
            raise ValueError(f"expected [B,3,{cfg.image},{cfg.image}], got {tuple(image.shape)}")
        kind = 1 if image.dtype == torch.uint8 else 0
        x = image.contiguous() if kind == 1 else image.to(torch.float32).contiguous()
        B = x.shape[0]
        out = torch.empty((B, cfg.out_dim), dtype=torch.float32, device=x.device)
        if B:
            _lib.check(_lib.load().drag_vit_encode(self._h, _lib.ptr(x), kind, B, _lib.ptr(out), int(normalize),
                                                   _lib.current_stream_ptr(x.device)), "drag_vit_encode")
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.load().drag_vit_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


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


def _transform_u8(n_px: int):
    """The host half of clip._transform only: Resize(bicubic, short side) -> CenterCrop -> RGB -> uint8 CHW. ToTensor and
    Normalize run on the GPU inside encode_image (same fp32 formula, so the embeddings equal the float path's)."""
    from torchvision import transforms as T
    return T.Compose([T.Resize(n_px, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(n_px),
                      lambda im: im.convert("RGB"), T.PILToTensor()])


def preprocess_u8(model: "CLIP"):
    """uint8 ingest transform for `model`: PIL -> uint8 [3,R,R]; stack, pin, one H2D per batch, encode_image."""
    return _transform_u8(model.visual.cfg.image)


def load(name: str = "ViT-B/32", device="cuda", state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 2000,
         max_batch: int = 512):
    """(model, preprocess) like clip.load(name, device). `state_dict` takes OpenAI `visual.*` keys."""
    if name not in CONFIGS:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    cfg = CONFIGS[name]
    state = state_dict if state_dict is not None else random_state(cfg, seed)
    return CLIP(CLIPVisual(cfg, state, device, max_batch=max_batch)), _transform(cfg.image)
