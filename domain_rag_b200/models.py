"""`load_model()` of the two generation scripts (batch_generate_flux_kshot.py:117-153,
outpainting_updown_sampling_redux.py:500-543): build the Redux prior pipeline and the Flux / Flux-Fill pipeline on
one GPU. Unlike the reference (which reloads 60 GB of weights PER SAMPLE, outpainting...:1185) the pipelines are
built once per process and shared (the SigLIP tower, Redux embedder, VAE and text table are common to both).

Weights: `weights_dir` with {flux_fill,flux_dev} / vae / siglip / redux as .pt (weights_only=True) or .safetensors, and
text_embeds.pt (scripts/make_text_embeds.py). The two transformers may be given in this package's fused layout or as a
diffusers FluxTransformer2DModel state dict (detected by its `x_embedder.weight` key and converted with
flux.from_diffusers_state_dict). A missing file is an ERROR (FileNotFoundError naming every missing file) - a run that
silently composed images from noise models would exit 0 with "completed" PNGs. Only `allow_random_init=True` (tests,
bench.py, the CLIs' --allow_random_init dry-run flag) substitutes seeded random weights and synthetic text tokens, and
says so for every component. `size="tiny"` builds reduced models for tests and CI.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import flux as F
from . import redux as R
from . import siglip as S
from .vae import FluxVAE


@dataclass
class Pipelines:
    prior_redux: R.FluxPriorReduxPipeline
    pipe: Optional[F.FluxPipeline]              # FLUX.1-dev shaped (in_channels 64) - background generation
    pipe_fill: Optional[F.FluxFillPipeline]     # FLUX.1-Fill-dev shaped (in_channels 384) - composition


def _maybe_load(weights_dir: Optional[str], name: str):
    """<weights_dir>/<name> (.pt, weights_only) or the same stem with .safetensors; None when neither exists."""
    if weights_dir:
        p = os.path.join(weights_dir, name)
        if os.path.exists(p):
            return torch.load(p, map_location="cpu", weights_only=True)
        st = os.path.splitext(p)[0] + ".safetensors"
        if os.path.exists(st):
            from safetensors.torch import load_file
            return load_file(st, device="cpu")
    return None


def required_files(want=("dev", "fill")):
    names = ["siglip.pt", "redux.pt", "text_embeds.pt", "vae.pt"]
    names += [f for kind, f in (("dev", "flux_dev.pt"), ("fill", "flux_fill.pt")) if kind in want]
    return names


def missing_files(weights_dir: Optional[str], want=("dev", "fill")):
    def present(name):
        if not weights_dir:
            return False
        p = os.path.join(weights_dir, name)
        return os.path.exists(p) or os.path.exists(os.path.splitext(p)[0] + ".safetensors")
    return [n for n in required_files(want) if not present(n)]


def _vae_random(seed: int, ch: int):
    # same seeded init as the oracle uses in the parity tests (oracle/vae.py::init_params), restated here because the
    # product never imports oracle/
    from .vae_init import init_params
    return init_params(seed=seed, ch=ch)


def load_model(device="cuda", want=("dev", "fill"), weights_dir: Optional[str] = None, size: str = "full",
               max_side: int = 1024, seed: int = 3000, max_batch: int = 1, allow_random_init: bool = False) -> Pipelines:
    dev = torch.device(device)
    tiny = size == "tiny"
    print("正在加载模型...")
    lacking = missing_files(weights_dir, want)
    if lacking and not allow_random_init:
        raise FileNotFoundError(
            f"load_model: missing weight files in {weights_dir!r}: {lacking}. The reference loads FLUX.1-dev / Fill-dev, the "
            "Redux prior (SigLIP + embedder), T5-XXL and CLIP-text here (batch_generate_flux_kshot.py:117-153, "
            "outpainting_updown_sampling_redux.py:500-543). Provide them (diffusers or fused layout; text_embeds.pt from "
            "scripts/make_text_embeds.py) or pass allow_random_init=True / --allow_random_init for a dry run with noise models")
    for name in lacking:
        print(f"警告: 未找到{name}, 使用随机初始化 (allow_random_init; 输出图像没有语义意义)")
    # --- image prompt path
    scfg = S.SiglipConfig(hidden=160, layers=2, heads=2, mlp=272, patch=14, image=60) if tiny else S.SiglipConfig()
    sstate = _maybe_load(weights_dir, "siglip.pt")
    rstate = _maybe_load(weights_dir, "redux.pt")
    txt_dim, pooled_dim, t5_tokens = (64, 32, 24) if tiny else (4096, 768, 512)
    if sstate is None:
        from .vae_init import init_siglip
        sstate = init_siglip(scfg, seed + 1)
    if rstate is None:
        from .vae_init import init_redux
        rstate = init_redux(seed + 2, scfg.hidden, 192 if tiny else 3 * 4096, txt_dim)
    have_text = bool(weights_dir) and os.path.exists(os.path.join(weights_dir, "text_embeds.pt"))
    table = R.TextEmbeddingTable(dev, txt_dim=txt_dim, pooled_dim=pooled_dim, tokens=t5_tokens,
                                 allow_synthetic=allow_random_init and not have_text)
    if have_text:
        table.load_file(os.path.join(weights_dir, "text_embeds.pt"))
    prior = R.FluxPriorReduxPipeline(S.SiglipVisionTower(scfg, sstate, dev), S.ReduxImageEncoder(rstate, dev), table)
    n_img_tokens = prior.image_encoder.cfg.tokens
    # --- VAE
    vstate = _maybe_load(weights_dir, "vae.pt")
    if vstate is None:
        vstate = _vae_random(seed + 3, 64 if tiny else 128)
    vae = FluxVAE(vstate, dev)
    # --- transformers
    base = dict(d=256, heads=2, n_double=2, n_single=2, txt_dim=txt_dim, pooled_dim=pooled_dim) if tiny else {}
    max_tokens = (max_side // 16) ** 2
    s_txt = t5_tokens + n_img_tokens
    pipes = {}
    for kind, cin, fname in (("dev", 64, "flux_dev.pt"), ("fill", 384, "flux_fill.pt")):
        if kind not in want:
            continue
        cfg = F.FluxConfig(in_channels=cin, **base)
        params = _maybe_load(weights_dir, fname)
        if params is None:
            params = F.init_params_device(cfg, seed=seed + (10 if kind == "dev" else 20), device=dev)
        elif "x_embedder.weight" in params:          # a diffusers FluxTransformer2DModel checkpoint
            params = F.from_diffusers_state_dict(params, cfg)
        tr = F.FluxTransformer(cfg, params, max_batch=max_batch, max_img_tokens=max_tokens, txt_tokens=s_txt, device=dev)
        pipes[kind] = F.FluxPipeline(tr, vae) if kind == "dev" else F.FluxFillPipeline(tr, vae)
    return Pipelines(prior, pipes.get("dev"), pipes.get("fill"))
