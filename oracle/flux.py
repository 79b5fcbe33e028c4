"""Oracle: Flux MMDiT forward, flow-match Euler sampler, latent packing, Redux blend (PyTorch fp32,
CPU). TEST INFRASTRUCTURE ONLY.

What the reference runs at batch_generate_flux_kshot.py:459-474 (FluxPriorReduxPipeline +
FluxPipeline) and outpainting_updown_sampling_redux.py:1237-1257 (FluxFillPipeline) is
diffusers==0.33.1 code that is NOT in /root/reference and not installable offline: PARITY UNPINNED
by the reference. This file restates the published architecture (FLUX.1-dev as converted by
diffusers' FluxTransformer2DModel; op order in SURVEY.md 2.3) and is cross-checked in
tests/test_flux_oracle.py against the independent BFL-style implementation shipped in this image
(torchtitan.experiments.flux.model) with a weight remap.

Parameter layout (shared with domain_rag_b200.flux - fused the way the kernels consume it):
  x_in / ctx_in                 Linear(C_in->d), Linear(4096->d)
  t_in / g_in / p_in            MLP embedders (Linear -> SiLU -> Linear) for timestep, guidance, pooled
  mod.w [n_mod, d], mod.b       every block's AdaLN modulation Linear stacked: per double block
                                img (shift1,scale1,gate1,shift2,scale2,gate2) then txt (same);
                                per single block (shift,scale,gate); final layer (scale,shift)
  double.i.{img,txt}.*          qkv [3d,d] (q rows, k rows, v rows), qnorm/knorm [128], out, mlp1, mlp2
  single.i.*                    qkv, qnorm/knorm, mlp [4d,d], out [d, 5d] (attn columns first)
  final.w [64, d], final.b
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass
class FluxConfig:
    in_channels: int = 64          # 64 FLUX.1-dev, 384 FLUX.1-Fill-dev
    d: int = 3072
    heads: int = 24                # head dim is 128 = sum(axes_dim)
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

    def mod_offset_double(self, i: int, txt: bool) -> int:
        return (i * 12 + (6 if txt else 0)) * self.d

    def mod_offset_single(self, i: int) -> int:
        return self.n_double * 12 * self.d + i * 3 * self.d

    def mod_offset_final(self) -> int:
        return self.n_double * 12 * self.d + self.n_single * 3 * self.d


def param_shapes(cfg: FluxConfig) -> Dict[str, tuple]:
    d, hd = cfg.d, 128
    s = {"x_in.w": (d, cfg.in_channels), "x_in.b": (d,), "ctx_in.w": (d, cfg.txt_dim), "ctx_in.b": (d,),
         "t_in.w1": (d, 256), "t_in.b1": (d,), "t_in.w2": (d, d), "t_in.b2": (d,),
         "p_in.w1": (d, cfg.pooled_dim), "p_in.b1": (d,), "p_in.w2": (d, d), "p_in.b2": (d,),
         "mod.w": (cfg.n_mod, d), "mod.b": (cfg.n_mod,),
         "final.w": (cfg.out_channels, d), "final.b": (cfg.out_channels,)}
    if cfg.guidance:
        s.update({"g_in.w1": (d, 256), "g_in.b1": (d,), "g_in.w2": (d, d), "g_in.b2": (d,)})
    for i in range(cfg.n_double):
        for st in ("img", "txt"):
            p = f"double.{i}.{st}."
            s.update({p + "qkv.w": (3 * d, d), p + "qkv.b": (3 * d,), p + "qnorm": (hd,), p + "knorm": (hd,),
                      p + "out.w": (d, d), p + "out.b": (d,),
                      p + "mlp1.w": (cfg.mlp_ratio * d, d), p + "mlp1.b": (cfg.mlp_ratio * d,),
                      p + "mlp2.w": (d, cfg.mlp_ratio * d), p + "mlp2.b": (d,)})
    for i in range(cfg.n_single):
        p = f"single.{i}."
        s.update({p + "qkv.w": (3 * d, d), p + "qkv.b": (3 * d,), p + "qnorm": (hd,), p + "knorm": (hd,),
                  p + "mlp.w": (cfg.mlp_ratio * d, d), p + "mlp.b": (cfg.mlp_ratio * d,),
                  p + "out.w": (d, (1 + cfg.mlp_ratio) * d), p + "out.b": (d,)})
    return s


def init_params(cfg: FluxConfig, seed: int = 3000, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Variance-preserving random init (no checkpoints offline, SURVEY 7 'hard parts'): Linear
    weights std = fan_in^-1/2, biases N(0, 0.02), modulation Linear small but non-zero so that
    shift/scale/gate paths are exercised, RMSNorm weights ~ 1."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("norm"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif name == "mod.w":
            t = torch.randn(shape, generator=g, device=device) * (0.5 * shape[1] ** -0.5)
        elif name == "mod.b":
            t = torch.randn(shape, generator=g, device=device) * 0.1
            # gates ~ N(0.5, 0.1) so residual branches contribute
        elif len(shape) == 2:
            t = torch.randn(shape, generator=g, device=device) * (shape[1] ** -0.5)
        else:
            t = torch.randn(shape, generator=g, device=device) * 0.02
        out[name] = t.to(dtype)
    return out


# ------------------------------------------------------------------------------------------ pieces
def timestep_embedding(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """[cos | sin] of 1000*t * exp(-ln(1e4) * i / half) (diffusers Timesteps with flip_sin_to_cos,
    downscale_freq_shift=0; same as torchtitan layers.py:35-59)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = (1000.0 * t.float())[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def rope_tables(ids: torch.Tensor, axes_dim=(16, 56, 56), theta: float = 10000.0):
    """ids [S,3] -> cos, sin fp32 [S, 64]; per axis a: omega_j = theta^(-2j/a), angle = pos*omega_j
    (computed in float64 like diffusers' get_1d_rotary_pos_embed with freqs_dtype=float64)."""
    cos, sin = [], []
    for i, a in enumerate(axes_dim):
        scale = torch.arange(0, a, 2, dtype=torch.float64) / a
        omega = 1.0 / (theta ** scale)
        ang = ids[:, i].double()[:, None] * omega[None]
        cos.append(torch.cos(ang))
        sin.append(torch.sin(ang))
    return torch.cat(cos, -1).float(), torch.cat(sin, -1).float()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B,H,S,128]; rotate interleaved pairs (2j, 2j+1) by angle j. Computed in fp32 and cast back to x.dtype
    (diffusers apply_rotary_emb: `.float()` ... `.type_as(query)`); a no-op cast for the fp32 oracle."""
    xf = x.float()
    x0, x1 = xf[..., 0::2], xf[..., 1::2]
    c, s = cos[None, None], sin[None, None]
    return torch.stack([x0 * c - x1 * s, x1 * c + x0 * s], dim=-1).flatten(-2).to(x.dtype)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Variance in fp32 (diffusers RMSNorm upcasts), result in x.dtype."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype) * w


def layer_norm(x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def _mlp_embed(p, name, x):
    return F.linear(F.silu(F.linear(x, p[name + ".w1"], p[name + ".b1"])), p[name + ".w2"], p[name + ".b2"])


def _qkv(p, prefix, h, heads):
    B, S, _ = h.shape
    qkv = F.linear(h, p[prefix + "qkv.w"], p[prefix + "qkv.b"]).view(B, S, 3, heads, 128)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    q = rms_norm(q, p[prefix + "qnorm"]).permute(0, 2, 1, 3)
    k = rms_norm(k, p[prefix + "knorm"]).permute(0, 2, 1, 3)
    return q, k, v.permute(0, 2, 1, 3)


def attention(q, k, v, cos, sin):
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    o = F.scaled_dot_product_attention(q, k, v)      # softmax(q k^T / sqrt(128)) v, no mask
    B, H, S, D = o.shape
    return o.permute(0, 2, 1, 3).reshape(B, S, H * D)


def temb_vector(p, cfg: FluxConfig, t, g, pooled):
    dt = pooled.dtype                       # sinusoids are built in fp32 and cast to the model dtype (diffusers)
    vec = _mlp_embed(p, "t_in", timestep_embedding(t).to(dt))
    if cfg.guidance:
        vec = vec + _mlp_embed(p, "g_in", timestep_embedding(g).to(dt))
    return vec + _mlp_embed(p, "p_in", pooled)


def flux_forward(p: Dict[str, torch.Tensor], cfg: FluxConfig, x, ctx, pooled, t, g, img_ids, txt_ids):
    """x [B,S_img,C_in], ctx [B,S_txt,txt_dim], pooled [B,pooled_dim], t/g [B] -> v [B,S_img,out]."""
    d, H = cfg.d, cfg.heads
    img = F.linear(x, p["x_in.w"], p["x_in.b"])
    txt = F.linear(ctx, p["ctx_in.w"], p["ctx_in.b"])
    vec = temb_vector(p, cfg, t, g, pooled)
    mod = F.linear(F.silu(vec), p["mod.w"], p["mod.b"])            # [B, n_mod]
    cos, sin = (a.to(x.device) for a in rope_tables(torch.cat([txt_ids, img_ids], 0).cpu(), cfg.axes_dim, cfg.theta))
    S_txt = txt.shape[1]

    def chunks(off, n):
        return [mod[:, off + j * d: off + (j + 1) * d][:, None, :] for j in range(n)]

    for i in range(cfg.n_double):
        ish1, isc1, ig1, ish2, isc2, ig2 = chunks(cfg.mod_offset_double(i, False), 6)
        tsh1, tsc1, tg1, tsh2, tsc2, tg2 = chunks(cfg.mod_offset_double(i, True), 6)
        pi, pt = f"double.{i}.img.", f"double.{i}.txt."
        iq, ik, iv = _qkv(p, pi, layer_norm(img) * (1 + isc1) + ish1, H)
        tq, tk, tv = _qkv(p, pt, layer_norm(txt) * (1 + tsc1) + tsh1, H)
        attn = attention(torch.cat([tq, iq], 2), torch.cat([tk, ik], 2), torch.cat([tv, iv], 2), cos, sin)
        t_attn, i_attn = attn[:, :S_txt], attn[:, S_txt:]
        img = img + ig1 * F.linear(i_attn, p[pi + "out.w"], p[pi + "out.b"])
        h = layer_norm(img) * (1 + isc2) + ish2
        img = img + ig2 * F.linear(F.gelu(F.linear(h, p[pi + "mlp1.w"], p[pi + "mlp1.b"]), approximate="tanh"),
                                   p[pi + "mlp2.w"], p[pi + "mlp2.b"])
        txt = txt + tg1 * F.linear(t_attn, p[pt + "out.w"], p[pt + "out.b"])
        h = layer_norm(txt) * (1 + tsc2) + tsh2
        txt = txt + tg2 * F.linear(F.gelu(F.linear(h, p[pt + "mlp1.w"], p[pt + "mlp1.b"]), approximate="tanh"),
                                   p[pt + "mlp2.w"], p[pt + "mlp2.b"])
    z = torch.cat([txt, img], 1)
    for i in range(cfg.n_single):
        sh, sc, gt = chunks(cfg.mod_offset_single(i), 3)
        ps = f"single.{i}."
        h = layer_norm(z) * (1 + sc) + sh
        q, k, v = _qkv(p, ps, h, H)
        attn = attention(q, k, v, cos, sin)
        mlp = F.gelu(F.linear(h, p[ps + "mlp.w"], p[ps + "mlp.b"]), approximate="tanh")
        z = z + gt * F.linear(torch.cat([attn, mlp], 2), p[ps + "out.w"], p[ps + "out.b"])
    img = z[:, S_txt:]
    scale, shift = chunks(cfg.mod_offset_final(), 2)                  # diffusers order: (scale, shift)
    img = layer_norm(img) * (1 + scale) + shift
    return F.linear(img, p["final.w"], p["final.b"])


# ------------------------------------------------------------------------------------ pipeline glue
def pack_latents(z: torch.Tensor) -> torch.Tensor:
    """[B,16,h,w] -> [B,(h/2)(w/2),64]  (b c (h ph) (w pw) -> b (h w) (c ph pw))."""
    B, C, h, w = z.shape
    return z.view(B, C, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (h // 2) * (w // 2), C * 4)


def unpack_latents(x: torch.Tensor, h: int, w: int) -> torch.Tensor:
    B, _, ch = x.shape
    C = ch // 4
    return x.view(B, h // 2, w // 2, C, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, C, h, w)


def image_ids(h2: int, w2: int) -> torch.Tensor:
    """[(h2*w2), 3]: channel 0 = 0, 1 = row, 2 = column (latent grid after 2x2 packing)."""
    ids = torch.zeros(h2, w2, 3)
    ids[..., 1] = torch.arange(h2)[:, None]
    ids[..., 2] = torch.arange(w2)[None, :]
    return ids.reshape(h2 * w2, 3)


def flow_match_sigmas(num_steps: int, seq_len: int) -> torch.Tensor:
    """diffusers FlowMatchEulerDiscreteScheduler with dynamic shifting as FluxPipeline drives it:
    sigma_i = linspace(1, 1/T, T); mu = lerp(seq_len; 256 -> 0.5, 4096 -> 1.15);
    sigma' = e^mu / (e^mu + (1/sigma - 1)); append 0.  Returns float32 [T+1]."""
    sig = torch.linspace(1.0, 1.0 / num_steps, num_steps, dtype=torch.float64)
    m = (1.15 - 0.5) / (4096 - 256)
    mu = seq_len * m + (0.5 - m * 256)
    sig = math.exp(mu) / (math.exp(mu) + (1.0 / sig - 1.0))
    return torch.cat([sig, torch.zeros(1, dtype=torch.float64)]).float()


def euler_step(x: torch.Tensor, v: torch.Tensor, sigma: float, sigma_next: float) -> torch.Tensor:
    """x <- (x.float() + (sigma_next - sigma) * v.float()).to(x.dtype)."""
    return (x.float() + (sigma_next - sigma) * v.float()).to(x.dtype)


def redux_blend(txt_tokens, img_tokens, pooled, s_embed, s_pool):
    """FluxPriorReduxPipeline output (reference batch_generate_flux_kshot.py:459-465):
    prompt_embeds = cat([T5 tokens [B,512,4096], redux tokens [B,729,4096]], 1) * s_embed[:,None,None];
    pooled * s_pool[:,None]; both summed over the image batch (keepdim)."""
    pe = torch.cat([txt_tokens, img_tokens], dim=1) * s_embed[:, None, None]
    pp = pooled * s_pool[:, None]
    return pe.sum(0, keepdim=True), pp.sum(0, keepdim=True)


def sample(p, cfg: FluxConfig, latents_packed, ctx, pooled, guidance: float, num_steps: int, h2: int, w2: int,
           extra_cond=None, start_step: int = 0):
    """Guidance-distilled flow-matching sampling loop (no CFG second pass). extra_cond [B,S,C'] is the
    Fill conditioning (masked-image latents + mask) concatenated on the channel axis every step."""
    B = latents_packed.shape[0]
    sig = flow_match_sigmas(num_steps, latents_packed.shape[1])
    img_ids, txt_ids = image_ids(h2, w2), torch.zeros(ctx.shape[1], 3)
    x = latents_packed
    dev = x.device
    g = torch.full((B,), guidance, device=dev)
    for i in range(start_step, num_steps):
        t = torch.full((B,), float(sig[i]), device=dev)
        inp = x if extra_cond is None else torch.cat([x, extra_cond], dim=-1)
        v = flux_forward(p, cfg, inp, ctx, pooled, t, g, img_ids, txt_ids)
        x = euler_step(x, v, float(sig[i]), float(sig[i + 1]))
    return x
