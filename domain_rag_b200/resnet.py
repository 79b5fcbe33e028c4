"""ResNet-50 stem "style" encoder - drop-in for ResNetEncoder + calc_mean_std of the reference
(retrieval/clip100_resnet_style_all_shots.py:51-74) and its call pattern (:197-200):

    features = model(img)                 # img float32 [B,3,256,256] in [0,1], on the GPU
    mean, std = calc_mean_std(features)   # each [B,64,1,1]
    feat = torch.cat([mean.squeeze(), std.squeeze()])

Here `model(img)` launches one fused sm_100a kernel (conv7x7/2 as a split-bf16 implicit GEMM on the tcgen05 tensor
cores + folded eval BatchNorm + ReLU + maxpool3x3/2 + per-channel moments) and returns a `StemStats` handle;
`calc_mean_std` unpacks it. The 64x64x64 feature map is never materialised. `img` may also be the raw uint8 pixels
(the / 255 then runs in the kernel). No CPU path.
"""
from __future__ import annotations

from typing import Mapping, Optional

import torch

from . import _lib

STYLE_EPS = 1e-5  # calc_mean_std eps (reference :68)
BN_EPS = 1e-5     # torchvision BatchNorm2d default


class StemStats:
    """Result of ResNetEncoder.__call__: per-image channel statistics of the stem output."""

    def __init__(self, stats: torch.Tensor):
        self.stats = stats  # [B, 128] = cat(mean, std)

    def size(self):
        b = self.stats.shape[0]
        return torch.Size((b, 64, 64, 64))


def calc_mean_std(feat, eps: float = STYLE_EPS):
    """Same signature/return as the reference helper; accepts the fused StemStats handle."""
    if not isinstance(feat, StemStats):
        raise TypeError("calc_mean_std expects the StemStats returned by ResNetEncoder(img)")
    if abs(eps - STYLE_EPS) > 0:
        raise ValueError("the fused kernel bakes eps=1e-5 (reference default) into the statistics")
    b = feat.stats.shape[0]
    mean = feat.stats[:, :64].reshape(b, 64, 1, 1)
    std = feat.stats[:, 64:].reshape(b, 64, 1, 1)
    return mean, std


def fold_stem(conv_w: torch.Tensor, bn_w: torch.Tensor, bn_b: torch.Tensor, bn_mean: torch.Tensor,
              bn_var: torch.Tensor, bn_eps: float = BN_EPS):
    """Fold eval-mode BatchNorm into the 7x7 filter: returns (w_fold [64,3,7,7], b_fold [64]) fp32."""
    scale = bn_w.double() / torch.sqrt(bn_var.double() + bn_eps)
    w = (conv_w.double() * scale[:, None, None, None]).float().contiguous()
    b = (bn_b.double() - bn_mean.double() * scale).float().contiguous()
    return w, b


def random_stem_state(seed: int = 2000) -> dict:
    """Synthetic stand-in for torchvision's pretrained conv1/bn1 (no checkpoints offline): conv
    std = fan_in^-1/2, BN gamma=1 beta=0, running_mean~U(-.1,.1), running_var~U(.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    return {
        "conv1.weight": torch.randn(64, 3, 7, 7, generator=g) * (147 ** -0.5),
        "bn1.weight": torch.ones(64),
        "bn1.bias": torch.zeros(64),
        "bn1.running_mean": torch.rand(64, generator=g) * 0.2 - 0.1,
        "bn1.running_var": torch.rand(64, generator=g) + 0.5,
    }


class ResNetEncoder:
    """conv1 -> bn1 -> relu -> maxpool of ResNet-50, fused with the style statistics."""

    def __init__(self, state: Optional[Mapping[str, torch.Tensor]] = None, seed: int = 2000):
        self.state = dict(state) if state is not None else random_stem_state(seed)
        self.w_fold, self.b_fold = fold_stem(self.state["conv1.weight"], self.state["bn1.weight"],
                                             self.state["bn1.bias"], self.state["bn1.running_mean"],
                                             self.state["bn1.running_var"])
        self.device = torch.device("cpu")

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ResNetEncoder runs only on CUDA (sm_100a); there is no CPU path")
        self.device = device
        self.w_fold = self.w_fold.to(device)
        self.b_fold = self.b_fold.to(device)
        return self

    def eval(self):
        return self

    def __call__(self, img: torch.Tensor) -> StemStats:
        if not img.is_cuda:
            raise RuntimeError("ResNetEncoder: input must be a CUDA tensor (no CPU path)")
        if self.w_fold.device != img.device:
            self.to(img.device)
        u8 = img.dtype == torch.uint8          # raw pixels: the / 255 of reference :193 runs inside the kernel
        img = img.contiguous() if u8 else img.float().contiguous()
        if img.dim() != 4 or img.shape[1] != 3:
            raise ValueError(f"expected [B,3,256,256], got {tuple(img.shape)}")
        b, _, h, w = img.shape
        out = torch.empty((b, 128), dtype=torch.float32, device=img.device)
        fn = _lib.load().drag_stem_stats_u8 if u8 else _lib.load().drag_stem_stats
        _lib.check(fn(_lib.ptr(img), b, h, w, _lib.ptr(self.w_fold), _lib.ptr(self.b_fold), STYLE_EPS, _lib.ptr(out),
                      _lib.current_stream_ptr(img.device)), "drag_stem_stats")
        return StemStats(out)

    def style_features(self, img: torch.Tensor) -> torch.Tensor:
        """[B,3,256,256] -> [B,128] = cat(mean, std): what compute_resnet_features returns per image."""
        return self(img).stats
