"""Oracle: ResNet-50 stem + style statistics (PyTorch fp32, CPU). TEST INFRASTRUCTURE ONLY.

Follows retrieval/clip100_resnet_style_all_shots.py: ResNetEncoder :51-64 (torchvision resnet50
conv1 -> bn1 -> relu -> maxpool, eval mode :228), calc_mean_std :67-74 (unbiased variance + 1e-5,
sqrt; mean), feature = cat(mean, std) :200, input scaling /255 without normalisation :191-193.
PINNED: tests/golden/stem_stats.npz holds outputs of the reference's own ResNetEncoder /
calc_mean_std / compute_resnet_features run through oracle/make_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def stem_forward(img: torch.Tensor, state: dict) -> torch.Tensor:
    """[B,3,H,W] fp32 -> [B,64,H/4,W/4] (conv1, eval bn1, relu, maxpool)."""
    x = F.conv2d(img.float(), state["conv1.weight"].float(), None, stride=2, padding=3)
    x = F.batch_norm(x, state["bn1.running_mean"].float(), state["bn1.running_var"].float(),
                     state["bn1.weight"].float(), state["bn1.bias"].float(), training=False, eps=1e-5)
    x = F.relu(x)
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


def calc_mean_std(feat: torch.Tensor, eps: float = 1e-5):
    n, c = feat.shape[:2]
    var = feat.view(n, c, -1).var(dim=2) + eps
    std = var.sqrt().view(n, c, 1, 1)
    mean = feat.view(n, c, -1).mean(dim=2).view(n, c, 1, 1)
    return mean, std


def style_features(img: torch.Tensor, state: dict) -> torch.Tensor:
    """[B,3,256,256] in [0,1] -> [B,128] = cat(mean64, std64)."""
    with torch.no_grad():
        mean, std = calc_mean_std(stem_forward(img, state))
        return torch.cat([mean.flatten(1), std.flatten(1)], dim=1)
