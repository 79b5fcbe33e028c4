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


def stem_forward(img: torch.Tensor, state: dict, dtype=torch.float32) -> torch.Tensor:
    """[B,3,H,W] -> [B,64,H/4,W/4] (conv1, eval bn1, relu, maxpool). dtype=float64 gives the
    high-precision comparator used for ill-conditioned (near-constant) inputs."""
    st = {k: v.to(dtype) for k, v in state.items()}
    x = F.conv2d(img.to(dtype), st["conv1.weight"], None, stride=2, padding=3)
    x = F.batch_norm(x, st["bn1.running_mean"], st["bn1.running_var"], st["bn1.weight"], st["bn1.bias"],
                     training=False, eps=1e-5)
    x = F.relu(x)
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


def calc_mean_std(feat: torch.Tensor, eps: float = 1e-5):
    n, c = feat.shape[:2]
    var = feat.view(n, c, -1).var(dim=2) + eps
    std = var.sqrt().view(n, c, 1, 1)
    mean = feat.view(n, c, -1).mean(dim=2).view(n, c, 1, 1)
    return mean, std


def style_features(img: torch.Tensor, state: dict, dtype=torch.float32) -> torch.Tensor:
    """[B,3,256,256] in [0,1] -> [B,128] = cat(mean64, std64)."""
    with torch.no_grad():
        mean, std = calc_mean_std(stem_forward(img, state, dtype))
        return torch.cat([mean.flatten(1), std.flatten(1)], dim=1).float()
