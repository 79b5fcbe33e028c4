"""Flux MMDiT on the sm_100a engine + the pipeline objects the reference scripts call.

Reference call sites:
  batch_generate_flux_kshot.py:467-474   pipe(guidance_scale=2.5, num_inference_steps=50, height=1024,
        width=1024, generator=torch.Generator("cpu").manual_seed(0), **pipe_prior_output).images
  outpainting_updown_sampling_redux.py:1246-1257   pipe_fill(image=..., mask_image=..., height=H, width=W,
        guidance_scale=g, num_inference_steps=50, prompt_embeds=..., pooled_prompt_embeds=...,
        generator=..., strength=s).images[0]
Both run diffusers' FluxTransformer2DModel once per step; here every step is one drag_flux_forward
(C++ orchestration over the tcgen05 GEMM / attention kernels) plus one drag_euler_step.
No PyTorch arithmetic on the step path; torch only owns the device buffers.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib


@dataclass
class FluxConfig:
    in_channels: int = 64          # 64 FLUX.1-dev (FluxPipeline), 384 FLUX.1-Fill-dev (FluxFillPipeline)
    d: int = 3072
    heads: int = 24
    n_double: int = 19
    n_single: int = 38
    txt_dim: int = 4096
    pooled_dim: int = 768
    out_channels: int = 64
    guidance: bool = True
    axes_dim: Tuple[int, int, int] = (16, 56, 56)
    theta: float = 10000.0
    mlp_ratio: int = 4

    @property
    def n_mod(self) -> int:
        return self.n_double * 12 * self.d + self.n_single * 3 * self.d + 2 * self.d


def param_shapes(cfg: FluxConfig) -> Dict[str, tuple]:
    """Fused parameter layout consumed by the engine (see oracle/flux.py for the meaning of each)."""
    d, hd, r = cfg.d, 128, cfg.mlp_ratio
    s = {"x_in.w": (d, cfg.in_channels), "x_in.b": (d,), "ctx_in.w": (d, cfg.txt_dim), "ctx_in.b": (d,),
         "t_in.w1": (d, 256), "t_in.b1": (d,), "t_in.w2": (d, d), "t_in.b2": (d,)}
    if cfg.guidance:
        s.update({"g_in.w1": (d, 256), "g_in.b1": (d,), "g_in.w2": (d, d), "g_in.b2": (d,)})
    s.update({"p_in.w1": (d, cfg.pooled_dim), "p_in.b1": (d,), "p_in.w2": (d, d), "p_in.b2": (d,),
              "mod.w": (cfg.n_mod, d), "mod.b": (cfg.n_mod,),
              "final.w": (cfg.out_channels, d), "final.b": (cfg.out_channels,)})
    for i in range(cfg.n_double):
        for st in ("img", "txt"):
            p = f"double.{i}.{st}."
            s.update({p + "qkv.w": (3 * d, d), p + "qkv.b": (3 * d,), p + "qnorm": (hd,), p + "knorm": (hd,),
                      p + "out.w": (d, d), p + "out.b": (d,), p + "mlp1.w": (r * d, d), p + "mlp1.b": (r * d,),
                      p + "mlp2.w": (d, r * d), p + "mlp2.b": (d,)})
    for i in range(cfg.n_single):
        p = f"single.{i}."
        s.update({p + "qkv.w": (3 * d, d), p + "qkv.b": (3 * d,), p + "qnorm": (hd,), p + "knorm": (hd,),
                  p + "mlp.w": (r * d, d), p + "mlp.b": (r * d,), p + "out.w": (d, (1 + r) * d), p + "out.b": (d,)})
    return s


def param_order(cfg: FluxConfig) -> List[Optional[str]]:
    """Canonical pointer order of drag_flux_set_weights: 20 globals (guidance slots None when the
    model has no guidance embedder), 20 per double block (img then txt), 8 per single block."""
    g = ["g_in.w1", "g_in.b1", "g_in.w2", "g_in.b2"] if cfg.guidance else [None] * 4
    order: List[Optional[str]] = ["x_in.w", "x_in.b", "ctx_in.w", "ctx_in.b", "t_in.w1", "t_in.b1", "t_in.w2",
                                  "t_in.b2", *g, "p_in.w1", "p_in.b1", "p_in.w2", "p_in.b2", "mod.w", "mod.b",
                                  "final.w", "final.b"]
    for i in range(cfg.n_double):
        for st in ("img", "txt"):
            p = f"double.{i}.{st}."
            order += [p + k for k in ("qkv.w", "qkv.b", "qnorm", "knorm", "out.w", "out.b", "mlp1.w", "mlp1.b",
                                      "mlp2.w", "mlp2.b")]
    for i in range(cfg.n_single):
        p = f"single.{i}."
        order += [p + k for k in ("qkv.w", "qkv.b", "qnorm", "knorm", "mlp.w", "mlp.b", "out.w", "out.b")]
    return order


def init_params_device(cfg: FluxConfig, seed: int = 3000, device="cuda") -> Dict[str, torch.Tensor]:
    """Random-init bf16 weights generated ON the device (11.9 B parameters at full size; no checkpoints
    offline). Variance preserving: Linear std = fan_in^-1/2, biases 0.02, modulation Linear half
    that (shift/scale/gate are exercised but stay O(1)), RMSNorm weights ~ 1. Large tensors are drawn
    in row chunks to bound the fp32 staging memory."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("norm"):
            t = (1.0 + 0.1 * torch.randn(shape, generator=g, device=device)).bfloat16()
        elif len(shape) == 2:
            std = shape[1] ** -0.5 * (0.5 if name == "mod.w" else 1.0)
            t = torch.empty(shape, dtype=torch.bfloat16, device=device)
            rows = max(1, (1 << 26) // shape[1])
            for r0 in range(0, shape[0], rows):
                r1 = min(shape[0], r0 + rows)
                t[r0:r1] = (torch.randn((r1 - r0, shape[1]), generator=g, device=device) * std).bfloat16()
        else:
            std = 0.1 if name == "mod.b" else 0.02
            t = (torch.randn(shape, generator=g, device=device) * std).bfloat16()
        out[name] = t
    return out


def rope_tables(ids: torch.Tensor, axes_dim=(16, 56, 56), theta: float = 10000.0):
    """ids [S,3] (host) -> (cos, sin) fp32 [S,64]; angles in float64 like diffusers' FluxPosEmbed."""
    cos, sin = [], []
    for i, a in enumerate(axes_dim):
        omega = 1.0 / (theta ** (torch.arange(0, a, 2, dtype=torch.float64) / a))
        ang = ids[:, i].double()[:, None] * omega[None]
        cos.append(torch.cos(ang))
        sin.append(torch.sin(ang))
    return torch.cat(cos, -1).float().contiguous(), torch.cat(sin, -1).float().contiguous()


def image_ids(h2: int, w2: int) -> torch.Tensor:
    ids = torch.zeros(h2, w2, 3)
    ids[..., 1] = torch.arange(h2)[:, None]
    ids[..., 2] = torch.arange(w2)[None, :]
    return ids.reshape(h2 * w2, 3)


def flow_match_sigmas(num_steps: int, seq_len: int) -> List[float]:
    """sigma_i = linspace(1, 1/T, T) shifted by mu(seq_len) (FlowMatchEulerDiscreteScheduler with
    dynamic shifting as the Flux pipelines configure it), with the terminal 0 appended."""
    m = (1.15 - 0.5) / (4096 - 256)
    mu = seq_len * m + (0.5 - m * 256)
    sig = torch.linspace(1.0, 1.0 / num_steps, num_steps, dtype=torch.float64)
    sig = math.exp(mu) / (math.exp(mu) + (1.0 / sig - 1.0))
    return [float(s) for s in sig.float()] + [0.0]


def pack_latents(z: torch.Tensor) -> torch.Tensor:
    """[B,16,h,w] -> [B,(h/2)(w/2),64] (data movement only)."""
    B, Cc, h, w = z.shape
    return z.view(B, Cc, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (h // 2) * (w // 2), Cc * 4)


def unpack_latents(x: torch.Tensor, h: int, w: int) -> torch.Tensor:
    B, _, ch = x.shape
    Cc = ch // 4
    return x.view(B, h // 2, w // 2, Cc, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, Cc, h, w)


def pack_latents_device(z: torch.Tensor, out: Optional[torch.Tensor] = None, ch_off: int = 0) -> torch.Tensor:
    """pack_latents of a bf16 CUDA tensor [B,C,h,w] as one kernel (drag_pack_latents); with `out` [B,S,ld] the packed
    channels land at out[..., ch_off : ch_off + 4C]."""
    B, Cc, h, w = z.shape
    z = z.contiguous()
    if out is None:
        out = torch.empty((B, (h // 2) * (w // 2), 4 * Cc), dtype=torch.bfloat16, device=z.device)
    assert z.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and out.stride(2) == 1
    assert out.stride(0) == out.shape[1] * out.stride(1)
    _lib.check(_lib.load().drag_pack_latents(_lib.ptr(z), B, Cc, h, w, _lib.ptr(out), out.stride(1), ch_off,
                                             _lib.current_stream_ptr(z.device)), "drag_pack_latents")
    return out


def unpack_latents_device(x: torch.Tensor, h: int, w: int, channels: int = 16) -> torch.Tensor:
    """unpack_latents of bf16 CUDA [B,S,ld >= 4*channels] (a channel-strided view is fine) -> [B,channels,h,w]."""
    B = x.shape[0]
    assert x.dtype == torch.bfloat16 and x.stride(2) == 1 and x.stride(0) == x.shape[1] * x.stride(1)
    z = torch.empty((B, channels, h, w), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().drag_unpack_latents(_lib.ptr(x), x.stride(1), B, channels, h, w, _lib.ptr(z),
                                               _lib.current_stream_ptr(x.device)), "drag_unpack_latents")
    return z


def pack_fill_inputs(latents: Optional[torch.Tensor], masked_latents: torch.Tensor, mask_u8: torch.Tensor) -> torch.Tensor:
    """The Flux-Fill transformer input [B,S,384] in one launch (drag_pack_fill_inputs): packed `latents` [B,S,64] (or None:
    channels 0:64 left for the caller), masked-image latents bf16 [B,16,h,w], mask uint8 [B,8h,8w] (non-zero = repaint)."""
    B, _, h, w = masked_latents.shape
    assert mask_u8.dtype == torch.uint8 and tuple(mask_u8.shape) == (B, 8 * h, 8 * w)
    x = torch.empty((B, (h // 2) * (w // 2), 384), dtype=torch.bfloat16, device=masked_latents.device)
    lat = latents.contiguous() if latents is not None else None
    _lib.check(_lib.load().drag_pack_fill_inputs(_lib.ptr(lat) if lat is not None else None, 64 if lat is not None else 0,
                                                 _lib.ptr(masked_latents.contiguous()), _lib.ptr(mask_u8.contiguous()), B, h, w,
                                                 _lib.ptr(x), 384, _lib.current_stream_ptr(x.device)),
               "drag_pack_fill_inputs")
    return x


class FluxTransformer:
    """FluxTransformer2DModel stand-in: owns bf16 device weights and one C++ engine."""

    def __init__(self, cfg: FluxConfig, params: Dict[str, torch.Tensor], max_batch: int, max_img_tokens: int,
                 txt_tokens: int, device="cuda"):
        lib = _lib.load()
        self.cfg, self.device = cfg, torch.device(device)
        self.params = {k: v.to(self.device, torch.bfloat16).contiguous() for k, v in params.items()}
        missing = [n for n in param_shapes(cfg) if n not in self.params]
        if missing:
            raise KeyError(f"missing Flux parameters: {missing[:4]}...")
        self.max_batch, self.max_img_tokens, self.txt_tokens = max_batch, max_img_tokens, txt_tokens
        c = (C.c_int * 12)(cfg.in_channels, cfg.d, cfg.heads, cfg.n_double, cfg.n_single, cfg.txt_dim, cfg.pooled_dim,
                           cfg.out_channels, int(cfg.guidance), max_batch, max_img_tokens, txt_tokens)
        self._h = C.c_void_p()
        _lib.check(lib.drag_flux_create(C.byref(c), C.byref(self._h)), "drag_flux_create")
        order = param_order(cfg)
        ptrs = (C.c_void_p * len(order))(*[self.params[n].data_ptr() if n else None for n in order])
        _lib.check(lib.drag_flux_set_weights(self._h, ptrs, len(order)), "drag_flux_set_weights")

    def forward(self, x, ctx, pooled, t, g, rope_cos, rope_sin, out=None, n_double_run=-1, n_single_run=-1):
        """x bf16 [B,S_img,ldx>=C_in] (may be a channel-strided view), ctx bf16 [B,S_txt,txt_dim], pooled bf16
        [B,pooled_dim], t/g fp32 [B] device tensors -> v bf16 [B,S_img,out_channels]."""
        B, S_img = x.shape[0], x.shape[1]
        assert x.stride(2) == 1 and x.stride(0) == S_img * x.stride(1)
        assert ctx.is_contiguous() and pooled.is_contiguous() and ctx.shape[1] == self.txt_tokens
        if out is None:
            out = torch.empty((B, S_img, self.cfg.out_channels), dtype=torch.bfloat16, device=x.device)
        _lib.check(_lib.load().drag_flux_forward(
            self._h, _lib.ptr(x), x.stride(1), _lib.ptr(ctx), _lib.ptr(pooled), _lib.ptr(t),
            _lib.ptr(g) if g is not None else None, _lib.ptr(rope_cos), _lib.ptr(rope_sin), B, S_img,
            _lib.ptr(out), out.stride(1), n_double_run, n_single_run, _lib.current_stream_ptr(x.device)),
            "drag_flux_forward")
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.load().drag_flux_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def euler_step_(x: torch.Tensor, v: torch.Tensor, dsigma: float) -> None:
    """In place x += dsigma * v on [B,S,C] bf16 views (channel-strided allowed)."""
    B, S, Cc = v.shape
    _lib.check(_lib.load().drag_euler_step(_lib.ptr(x), x.stride(1), _lib.ptr(v), v.stride(1), B * S, Cc,
                                           float(dsigma), _lib.current_stream_ptr(x.device)), "drag_euler_step")


def redux_blend(txt_tokens, img_tokens, pooled, s_embed, s_pool):
    """FluxPriorReduxPipeline output from precomputed encoder tokens: bf16 [B,512,4096], [B,729,4096],
    [B,768], fp32 scales [B] -> (prompt_embeds [1,1241,4096], pooled_prompt_embeds [1,768])."""
    B, n_txt, dim = txt_tokens.shape
    n_img = img_tokens.shape[1]
    dev = txt_tokens.device
    out_e = torch.empty((1, n_txt + n_img, dim), dtype=torch.bfloat16, device=dev)
    out_p = torch.empty((1, pooled.shape[1]), dtype=torch.bfloat16, device=dev)
    se = torch.as_tensor(s_embed, dtype=torch.float32, device=dev).contiguous()
    sp = torch.as_tensor(s_pool, dtype=torch.float32, device=dev).contiguous()
    _lib.check(_lib.load().drag_redux_blend(_lib.ptr(txt_tokens.contiguous()), _lib.ptr(img_tokens.contiguous()),
                                            _lib.ptr(pooled.contiguous()), _lib.ptr(se), _lib.ptr(sp), _lib.ptr(out_e),
                                            _lib.ptr(out_p), B, n_txt, n_img, dim, pooled.shape[1],
                                            _lib.current_stream_ptr(dev)), "drag_redux_blend")
    return out_e, out_p


@dataclass
class PipelineOutput:
    latents: torch.Tensor                      # [B,16,H/8,W/8] bf16 (unpacked, sampler space)
    images: Optional[list] = None              # PIL images when the pipeline owns a VAE and output_type == "pil"
    steps_run: int = 0


def axpby_(x: torch.Tensor, y: torch.Tensor, a: float, b: float) -> torch.Tensor:
    """a * x + b * y on contiguous bf16 tensors (drag_axpby_bf16)."""
    out = torch.empty_like(x)
    _lib.check(_lib.load().drag_axpby_bf16(_lib.ptr(x), _lib.ptr(y), float(a), float(b), _lib.ptr(out), x.numel(),
                                           _lib.current_stream_ptr(x.device)), "drag_axpby_bf16")
    return out


def executed_range(num_inference_steps: int, strength: float) -> int:
    """First executed step index of an img2img-style call, as diffusers' Flux get_timesteps computes it:
    t_start = int(max(T - min(T * strength, T), 0)) - the truncation happens AFTER the subtraction, so a fractional
    T * strength rounds the start DOWN (T = 50, strength 0.75 -> start 12, 38 executed steps)."""
    return int(max(num_inference_steps - min(num_inference_steps * strength, num_inference_steps), 0))


class FluxPipeline:
    """Mirror of the diffusers pipeline call the reference makes (kwargs, CPU generator semantics, schedule, Euler
    update): batch_generate_flux_kshot.py:467-474. Prompt tensors come from FluxPriorReduxPipeline (redux.py); with a
    `vae` the call returns `.images` (PIL), otherwise latents (output_type="latent")."""

    def __init__(self, transformer: FluxTransformer, vae=None):
        self.transformer = transformer
        self.vae = vae
        self._rope_cache = {}

    def _rope(self, h2, w2, s_txt, device):
        key = (h2, w2, s_txt)
        if key not in self._rope_cache:
            ids = torch.cat([torch.zeros(s_txt, 3), image_ids(h2, w2)], 0)
            cos, sin = rope_tables(ids, self.transformer.cfg.axes_dim, self.transformer.cfg.theta)
            self._rope_cache[key] = (cos.to(device), sin.to(device))
        return self._rope_cache[key]

    def prepare_latents(self, batch, height, width, generator, device):
        """randn([B,16,H/8,W/8]) drawn with the CPU generator in bf16, then moved to the GPU (the reference
        passes torch.Generator("cpu").manual_seed(0), batch_generate_flux_kshot.py:472); H, W floored to /16."""
        h, w = 2 * (int(height) // 16), 2 * (int(width) // 16)
        if isinstance(generator, (list, tuple)):          # one generator per batch element: each draws its own [1,16,h,w]
            if len(generator) != batch:
                raise ValueError(f"got {len(generator)} generators for a batch of {batch}")
            z = torch.cat([torch.randn((1, 16, h, w), generator=g, dtype=torch.bfloat16) for g in generator])
        else:
            z = torch.randn((batch, 16, h, w), generator=generator, dtype=torch.bfloat16)
        return pack_latents(z).contiguous().pin_memory().to(device, non_blocking=True), h, w

    def _denoise(self, latents, h, w, prompt_embeds, pooled_prompt_embeds, guidance_scale, num_inference_steps, start,
                 extra_cond, buf=None):
        """Steps [start, T) of the flow-match Euler loop on packed latents [B,S,64]; returns the packed latents as a view
        of the working buffer. `buf` (Fill: [B,S,384] from pack_fill_inputs, latents already in channels 0:64) is used as
        is; otherwise it is built from `latents` (+ `extra_cond`)."""
        tr = self.transformer
        dev = tr.device
        if buf is not None:
            B, S_img, c_lat = buf.shape[0], buf.shape[1], tr.cfg.out_channels
        else:
            B, S_img, c_lat = latents.shape
            if extra_cond is not None:   # x lives in the first 64 channels of a persistent [B,S,64+cond] buffer
                buf = torch.cat([latents, extra_cond.to(latents.dtype)], dim=-1).contiguous()
            else:
                buf = latents.contiguous()
        sig = flow_match_sigmas(num_inference_steps, S_img)
        cos, sin = self._rope(h // 2, w // 2, prompt_embeds.shape[1], dev)
        x_view = buf[:, :, :c_lat]
        ctx = prompt_embeds.to(dev, torch.bfloat16).contiguous()
        if ctx.shape[0] != B:
            ctx = ctx.expand(B, -1, -1).contiguous()
        pooled = pooled_prompt_embeds.to(dev, torch.bfloat16).contiguous()
        if pooled.shape[0] != B:
            pooled = pooled.expand(B, -1).contiguous()
        g = torch.full((B,), float(guidance_scale), dtype=torch.float32, device=dev) if tr.cfg.guidance else None
        t_all = torch.tensor(sig[:-1], dtype=torch.float32, device=dev)[:, None].expand(-1, B).contiguous()
        v = torch.empty((B, S_img, tr.cfg.out_channels), dtype=torch.bfloat16, device=dev)
        for i in range(start, num_inference_steps):
            tr.forward(buf, ctx, pooled, t_all[i], g, cos, sin, out=v)
            euler_step_(x_view, v, sig[i + 1] - sig[i])
        return x_view, sig

    def _finish(self, packed, h, w, steps_run, output_type):
        lat = unpack_latents_device(packed, h, w, packed.shape[2] // 4)
        if output_type == "latent":
            return PipelineOutput(latents=lat, images=None, steps_run=steps_run)
        if self.vae is None:
            raise RuntimeError("this pipeline was built without a VAE: pass vae=FluxVAE(...) or output_type='latent'")
        if output_type == "u8":          # uint8 [B,H,W,3] left on the device (bench.py's device-resident leg)
            return PipelineOutput(latents=lat, images=self.vae.decode(lat, output_type="u8"), steps_run=steps_run)
        return PipelineOutput(latents=lat, images=self.vae.decode(lat, output_type="pil"), steps_run=steps_run)

    def __call__(self, prompt_embeds, pooled_prompt_embeds, guidance_scale=3.5, num_inference_steps=28, height=1024,
                 width=1024, generator=None, latents=None, extra_cond=None, strength=1.0, image_latents=None,
                 output_type=None):
        dev = self.transformer.device
        output_type = output_type or ("pil" if self.vae is not None else "latent")
        B = prompt_embeds.shape[0]
        if latents is None:
            latents, h, w = self.prepare_latents(B, height, width, generator, dev)
        else:
            h, w = 2 * (int(height) // 16), 2 * (int(width) // 16)
        start = 0
        if strength < 1.0:           # img2img: run only the last int(T*strength) steps from a noised image
            start = executed_range(num_inference_steps, strength)
            if image_latents is not None:
                s0 = flow_match_sigmas(num_inference_steps, latents.shape[1])[start]
                latents = axpby_(latents.contiguous(), image_latents.to(dev, torch.bfloat16).contiguous(), s0, 1.0 - s0)
        packed, _ = self._denoise(latents, h, w, prompt_embeds, pooled_prompt_embeds, guidance_scale, num_inference_steps,
                                  start, extra_cond)
        return self._finish(packed, h, w, num_inference_steps - start, output_type)


def pack_mask(mask: torch.Tensor) -> torch.Tensor:
    """Binary mask [B,H,W] -> [B,(H/16)(W/16),256]: each latent pixel carries its 8x8 block of mask pixels as 64
    channels (FluxFillPipeline.prepare_mask_latents), then the 2x2 latent packing. Data movement only."""
    B, H, W = mask.shape
    m = mask.view(B, H // 8, 8, W // 8, 8).permute(0, 2, 4, 1, 3).reshape(B, 64, H // 8, W // 8)
    return pack_latents(m)


class FluxFillPipeline(FluxPipeline):
    """Mirror of the FluxFillPipeline call of the composition script (outpainting_updown_sampling_redux.py:1246-1257):
    pipe_fill(image=, mask_image=, height=, width=, guidance_scale=, num_inference_steps=50, prompt_embeds=,
    pooled_prompt_embeds=, generator=, strength=).images[0]. `image` / `mask_image` may also be equally long lists
    (one composition per entry, prompt tensors [B,...] or [1,...] broadcast; the C4 per-GPU slice runs its 4
    compositions as one batch). `generator` may be a list with one CPU generator per composition: every composition then
    draws exactly what a batch-1 call with that generator draws, so a batch reproduces the reference's per-composition
    seeds (outpainting...:1230-1231). The transformer is the Fill variant (in_channels 384 =
    64 latent + 64 masked-image latent + 256 mask channels). Host work: PIL resize to multiples of 16 (Lanczos, like
    VaeImageProcessor), mask binarisation at 0.5. Generator draws, in order: VAE sample of the image, the initial noise
    (bf16, CPU generator), VAE sample of the masked image."""

    def __call__(self, prompt_embeds, pooled_prompt_embeds, image=None, mask_image=None, height=None, width=None,
                 guidance_scale=30.0, num_inference_steps=50, generator=None, strength=1.0, output_type=None):
        import numpy as np
        from PIL import Image
        if self.vae is None:
            raise RuntimeError("FluxFillPipeline needs a VAE (image and masked image are encoded)")
        if image is None or mask_image is None:
            raise ValueError("image and mask_image are required")
        dev = self.transformer.device
        output_type = output_type or "pil"
        images = list(image) if isinstance(image, (list, tuple)) else [image]
        masks = list(mask_image) if isinstance(mask_image, (list, tuple)) else [mask_image]
        if len(masks) != len(images):
            raise ValueError("image and mask_image must have the same length")
        B = len(images)
        if B > self.transformer.max_batch:
            raise ValueError(f"batch {B} exceeds the transformer's max_batch {self.transformer.max_batch}")
        height = images[0].height if height is None else int(height)
        width = images[0].width if width is None else int(width)
        H, W = 16 * (height // 16), 16 * (width // 16)
        img_np, msk_np = [], []
        for im, mk in zip(images, masks):       # every image of a batch is resized to the same (W, H), like diffusers
            im = im.convert("RGB")
            if im.size != (W, H):
                im = im.resize((W, H), Image.LANCZOS)
            mk = mk.convert("L")
            if mk.size != (W, H):
                mk = mk.resize((W, H), Image.LANCZOS)
            img_np.append(np.asarray(im))
            msk_np.append((np.asarray(mk) >= 128).astype(np.uint8))
        img_u8 = torch.from_numpy(np.stack(img_np)).pin_memory().to(dev, non_blocking=True)
        mask_u8 = torch.from_numpy(np.stack(msk_np)).pin_memory().to(dev, non_blocking=True)
        return self.run_resident(img_u8, mask_u8, prompt_embeds, pooled_prompt_embeds, guidance_scale, num_inference_steps,
                                 strength, generator, output_type)

    def run_resident(self, img_u8, mask_u8, prompt_embeds, pooled_prompt_embeds, guidance_scale, num_inference_steps,
                     strength, generator, output_type="pil", noise=None):
        """Everything of __call__ after the host-side PIL work, on device-resident inputs: img_u8 uint8 [B,H,W,3],
        mask_u8 uint8 [B,H,W] (1 = repaint), H and W multiples of 16. `noise` (packed bf16 [B,S,64]) replaces the CPU
        generator's initial-noise draw (bench.py's `value` leg keeps it resident; `generator` may then be a device
        generator for the two VAE samples). __call__ and the bench share this one code path."""
        dev = self.transformer.device
        B, H, W = mask_u8.shape
        start = executed_range(num_inference_steps, strength)
        if start >= num_inference_steps:
            raise ValueError(f"After adjusting the num_inference_steps by strength parameter: {strength}, the number of "
                             "pipeline steps is 0 which is < 1")
        h, w = H // 8, W // 8
        image_latents = pack_latents_device(self.vae.encode(img_u8, generator=generator))
        if noise is None:
            noise, _, _ = self.prepare_latents(B, H, W, generator, dev)
        s0 = flow_match_sigmas(num_inference_steps, noise.shape[1])[start]
        latents = axpby_(noise.contiguous(), image_latents, s0, 1.0 - s0)
        masked = self.vae.encode(img_u8, generator=generator, mask=mask_u8)                          # [B,16,h,w]
        buf = pack_fill_inputs(latents, masked, mask_u8)                                             # [B,S,384]
        packed, _ = self._denoise(None, h, w, prompt_embeds, pooled_prompt_embeds, guidance_scale, num_inference_steps,
                                  start, None, buf=buf)
        return self._finish(packed, h, w, num_inference_steps - start, output_type)


def from_diffusers_state_dict(sd: Dict[str, torch.Tensor], cfg: FluxConfig) -> Dict[str, torch.Tensor]:
    """FluxTransformer2DModel (diffusers) checkpoint keys -> the fused layout of `param_shapes` (data movement only):
    q/k/v projections stacked into one [3d, d] matrix, the single blocks' proj_mlp kept separate from qkv, every
    AdaLN modulation Linear stacked into `mod.w` in block order (img then txt per double block; diffusers' chunk
    order shift/scale/gate and, for the final layer, scale/shift is already the order the engine reads)."""
    out: Dict[str, torch.Tensor] = {}

    def lin(dst_w, dst_b, src):
        out[dst_w], out[dst_b] = sd[src + ".weight"], sd[src + ".bias"]

    lin("x_in.w", "x_in.b", "x_embedder")
    lin("ctx_in.w", "ctx_in.b", "context_embedder")
    for dst, src in (("t_in", "timestep_embedder"), ("g_in", "guidance_embedder"), ("p_in", "text_embedder")):
        if dst == "g_in" and not cfg.guidance:
            continue
        lin(f"{dst}.w1", f"{dst}.b1", f"time_text_embed.{src}.linear_1")
        lin(f"{dst}.w2", f"{dst}.b2", f"time_text_embed.{src}.linear_2")
    mod_w, mod_b = [], []
    for i in range(cfg.n_double):
        b = f"transformer_blocks.{i}."
        for st, norm, q, k, v, o, nq, nk, ff in (
                ("img", "norm1", "attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0", "attn.norm_q", "attn.norm_k", "ff"),
                ("txt", "norm1_context", "attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj", "attn.to_add_out",
                 "attn.norm_added_q", "attn.norm_added_k", "ff_context")):
            p = f"double.{i}.{st}."
            mod_w.append(sd[b + norm + ".linear.weight"])
            mod_b.append(sd[b + norm + ".linear.bias"])
            out[p + "qkv.w"] = torch.cat([sd[b + n + ".weight"] for n in (q, k, v)], 0)
            out[p + "qkv.b"] = torch.cat([sd[b + n + ".bias"] for n in (q, k, v)], 0)
            out[p + "qnorm"], out[p + "knorm"] = sd[b + nq + ".weight"], sd[b + nk + ".weight"]
            lin(p + "out.w", p + "out.b", b + o)
            lin(p + "mlp1.w", p + "mlp1.b", b + ff + ".net.0.proj")
            lin(p + "mlp2.w", p + "mlp2.b", b + ff + ".net.2")
    for i in range(cfg.n_single):
        b, p = f"single_transformer_blocks.{i}.", f"single.{i}."
        mod_w.append(sd[b + "norm.linear.weight"])
        mod_b.append(sd[b + "norm.linear.bias"])
        out[p + "qkv.w"] = torch.cat([sd[b + f"attn.to_{n}.weight"] for n in "qkv"], 0)
        out[p + "qkv.b"] = torch.cat([sd[b + f"attn.to_{n}.bias"] for n in "qkv"], 0)
        out[p + "qnorm"], out[p + "knorm"] = sd[b + "attn.norm_q.weight"], sd[b + "attn.norm_k.weight"]
        lin(p + "mlp.w", p + "mlp.b", b + "proj_mlp")
        lin(p + "out.w", p + "out.b", b + "proj_out")
    mod_w.append(sd["norm_out.linear.weight"])
    mod_b.append(sd["norm_out.linear.bias"])
    out["mod.w"], out["mod.b"] = torch.cat(mod_w, 0), torch.cat(mod_b, 0)
    lin("final.w", "final.b", "proj_out")
    missing = [n for n in param_shapes(cfg) if n not in out]
    if missing:
        raise KeyError(f"checkpoint lacks parameters for: {missing[:4]}")
    for n, shape in param_shapes(cfg).items():
        if tuple(out[n].shape) != tuple(shape):
            raise ValueError(f"{n}: checkpoint shape {tuple(out[n].shape)} != expected {tuple(shape)}")
    return out
