"""Oracle: the three diffusers pipeline calls of the reference composed from the oracle modules (PyTorch fp32, CPU).
TEST INFRASTRUCTURE ONLY.

  redux_prior   FluxPriorReduxPipeline.__call__   batch_generate_flux_kshot.py:459-465, outpainting_...:1237-1243
  generate      FluxPipeline.__call__             batch_generate_flux_kshot.py:467-474
  fill          FluxFillPipeline.__call__         outpainting_updown_sampling_redux.py:1246-1257
diffusers==0.33.1 is not in /root/reference and not installable offline: PARITY UNPINNED by the reference; the
control flow below restates the published pipelines (preprocess -> VAE encode -> scale_noise -> mask packing ->
flow-match Euler loop over the steps kept by `strength` -> VAE decode -> postprocess).
"""
from __future__ import annotations

import numpy as np
import torch

from . import flux as OF
from . import siglip as OS
from . import vae as OV


def redux_prior(siglip_state, siglip_cfg, redux_state, images, txt_tokens, pooled, s_embed, s_pool):
    """images: PIL list; txt_tokens [B,512,D], pooled [B,P] (the constant text half). -> (prompt_embeds, pooled)."""
    px = OS.preprocess(images, siglip_cfg.image)
    img_tokens = OS.redux_embed(redux_state, OS.last_hidden_state(siglip_state, siglip_cfg, px))
    return OF.redux_blend(txt_tokens.float(), img_tokens, pooled.float(), torch.tensor(s_embed, dtype=torch.float32),
                          torch.tensor(s_pool, dtype=torch.float32))


def executed_start(num_steps: int, strength: float) -> int:
    """diffusers 0.33.1 FluxFillPipeline.get_timesteps: init = min(T * s, T); t_start = int(max(T - init, 0))."""
    return int(max(num_steps - min(num_steps * strength, num_steps), 0))


def pack_mask(mask: torch.Tensor) -> torch.Tensor:
    B, H, W = mask.shape
    m = mask.view(B, H // 8, 8, W // 8, 8).permute(0, 2, 4, 1, 3).reshape(B, 64, H // 8, W // 8)
    return OF.pack_latents(m)


def generate(p_flux, cfg, p_vae, prompt_embeds, pooled, guidance, num_steps, height, width, generator):
    h, w = 2 * (height // 16), 2 * (width // 16)
    z = torch.randn((1, 16, h, w), generator=generator, dtype=torch.bfloat16)
    x = OF.sample(p_flux, cfg, OF.pack_latents(z).float(), prompt_embeds.float(), pooled.float(), guidance, num_steps,
                  h // 2, w // 2)
    lat = OF.unpack_latents(x, h, w)
    return lat, OV.postprocess_u8(OV.decode_latents(lat, p_vae))


def fill(p_flux, cfg, p_vae, image_u8: np.ndarray, mask_bool: np.ndarray, prompt_embeds, pooled, guidance, num_steps,
         strength, generator, device="cpu", flux_dtype=torch.float32, vae_dtype=torch.float32):
    """image_u8 [H,W,3] or [B,H,W,3], mask_bool [H,W] or [B,H,W] (True = repaint), H and W multiples of 16.
    -> (latents, image uint8 [B,H,W,3]). A batch shares one generator: each draw covers the whole batch.
    device / flux_dtype: the torch-ref legs (oracle/torchref.py) run the same control flow on the GPU - random draws
    still come from the CPU generator, the transformer runs in `flux_dtype` on `p_flux` as given (a CastingParams view
    for bf16 device weights) and the VAE in `vae_dtype` (the reference pipeline is bf16 throughout, VAE included)."""
    if image_u8.ndim == 3:
        image_u8, mask_bool = image_u8[None], mask_bool[None]
    B, H, W = mask_bool.shape
    h, w = H // 8, W // 8
    dev = torch.device(device)
    if dev.type != "cpu" or vae_dtype != torch.float32:
        p_vae = {k: v.to(dev, vae_dtype) for k, v in p_vae.items()}
    img = OV.preprocess_image(torch.from_numpy(np.ascontiguousarray(image_u8))).to(dev, vae_dtype)
    mask = torch.from_numpy(mask_bool.astype(np.float32)).to(dev)
    prompt_embeds, pooled = prompt_embeds.to(dev), pooled.to(dev)
    if prompt_embeds.shape[0] != B:
        prompt_embeds, pooled = prompt_embeds.expand(B, -1, -1), pooled.expand(B, -1)
    start = executed_start(num_steps, strength)

    gens = list(generator) if isinstance(generator, (list, tuple)) else None   # one generator per composition

    def draw(shape):
        if gens is None:
            return torch.randn(tuple(shape), generator=generator, dtype=torch.bfloat16).to(dev)
        return torch.cat([torch.randn((1,) + tuple(shape[1:]), generator=g, dtype=torch.bfloat16) for g in gens]).to(dev)

    def vae_sample(x):
        mean, logvar = OV.encoder(x.to(vae_dtype), p_vae).float().chunk(2, dim=1)
        noise = draw(mean.shape).float()
        return ((mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise) - OV.SHIFT_FACTOR) * OV.SCALE_FACTOR

    image_latents = OF.pack_latents(vae_sample(img))
    noise = OF.pack_latents(draw((B, 16, h, w))).float()
    s0 = float(OF.flow_match_sigmas(num_steps, noise.shape[1])[start])
    latents = s0 * noise + (1.0 - s0) * image_latents
    masked = OF.pack_latents(vae_sample(img * (1.0 - mask[:, None])))
    cond = torch.cat([masked, pack_mask(mask)], dim=-1)
    # the pipeline holds latents / conditioning in the transformer dtype (bf16 in the reference: the scale_noise blend and
    # the masked-image latents are rounded once here); a no-op for the fp32 oracle
    x = OF.sample(p_flux, cfg, latents.to(flux_dtype), prompt_embeds.to(flux_dtype), pooled.to(flux_dtype), guidance,
                  num_steps, h // 2, w // 2, extra_cond=cond.to(flux_dtype), start_step=start)
    lat = OF.unpack_latents(x, h, w)
    return lat, OV.postprocess_u8(OV.decode_latents(lat.to(vae_dtype), p_vae).float())
