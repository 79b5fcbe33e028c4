"""Flux VAE (diffusers AutoencoderKL of FLUX.1-dev / Fill-dev) on the sm_100a kernels: the decode that turns the
sampler's latents into `pipe(...).images` and the encode that `pipe_fill(image=..., mask_image=...)` applies to the
conditioning image (reference: batch_generate_flux_kshot.py:467-474, outpainting_updown_sampling_redux.py:1246-1257).

Layout: activations bf16 NHWC; every 3x3 / 1x1 convolution is an implicit GEMM on the tcgen05 core (A tiles are 4-D
TMA boxes of the activation, image borders = TMA zero fill, weights [C_out][tap][C_in]); GroupNorm+SiLU, nearest
upsampling and the layout / pixel conversions are HBM-bound kernels (csrc/vae_ops.cu). The single-head mid-block
attention (head dim 512) is four GEMMs and a row softmax: Q K^T in fp32, P V through a transposed V projection;
the V bias is folded into the output projection bias (softmax rows sum to 1) and 1/sqrt(C) into the Q projection.
No PyTorch arithmetic on the path and no CPU fallback; torch only owns the buffers. Parameter names follow
oracle/vae.py (`from_diffusers_state_dict` converts an AutoencoderKL checkpoint).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib, ops

SCALE_FACTOR = 0.3611
SHIFT_FACTOR = 0.1159
GN_GROUPS, GN_EPS = 32, 1e-6
EPI_BIAS, EPI_SILU, EPI_RESID, EPI_F32 = 0, 3, 4, 6
_ATTN_SCORE_BYTES = 1 << 30          # cap of the fp32 score block of the mid attention (rows are chunked)


def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class _Conv:
    """One convolution's device weights in the layout the implicit GEMM consumes."""

    def __init__(self, w: torch.Tensor, b: torch.Tensor, device, scale: float = 1.0):
        cout, cin, k, _ = w.shape
        self.k, self.cin, self.cout = k, _pad_to(cin, 64), _pad_to(cout, 64) if cout < 64 else cout
        wt = torch.zeros(self.cout, k * k, self.cin)
        wt[:cout, :, :cin] = (w.float() * scale).permute(0, 2, 3, 1).reshape(cout, k * k, cin)
        bias = torch.zeros(self.cout)
        bias[:cout] = b.float() * scale
        self.w = wt.to(device, torch.bfloat16).contiguous()
        self.b = bias.to(device, torch.bfloat16).contiguous()


class FluxVAE:
    """AutoencoderKL stand-in: `decode(latents)` / `encode(image)` with the pipelines' scale and shift applied."""

    def __init__(self, params: Dict[str, torch.Tensor], device="cuda"):
        _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FluxVAE runs only on CUDA (sm_100a); there is no CPU path")
        p = {k: v.detach().float().cpu() for k, v in params.items()}
        self.conv: Dict[str, _Conv] = {}
        self.norm: Dict[str, tuple] = {}
        self.lin: Dict[str, tuple] = {}
        for name in p:
            if name.endswith(".w") and p[name].dim() == 4 and ".attn." not in name:
                base = name[:-2]
                self.conv[base] = _Conv(p[name], p[base + ".b"], self.device)
            elif name.endswith(".w") and p[name].dim() == 1:
                base = name[:-2]
                if p[name].numel() % 64:
                    raise ValueError(f"{base}: {p[name].numel()} channels - the NHWC kernels need multiples of 64 "
                                     "(FLUX.1 VAE: 128/256/512)")
                self.norm[base] = (p[name].to(self.device, torch.bfloat16), p[base + ".b"].to(self.device, torch.bfloat16))
        for side in ("enc", "dec"):
            a = f"{side}.mid.attn"
            if a + ".q.w" not in p:
                continue
            c = p[a + ".q.w"].shape[0]
            bf = lambda t: t.to(self.device, torch.bfloat16).contiguous()  # noqa: E731
            wq, wk, wv, wo = (p[f"{a}.{n}.w"].reshape(c, c) for n in ("q", "k", "v", "proj"))
            s = c ** -0.5
            self.lin[a] = (bf(wq * s), bf(p[a + ".q.b"] * s), bf(wk), bf(p[a + ".k.b"]), bf(wv), bf(wo),
                           bf(wo @ p[a + ".v.b"] + p[a + ".proj.b"]))
        self.n_up = sum(1 for k in p if k.startswith("dec.up") and k.endswith(".res0.norm1.w"))
        self.n_down = sum(1 for k in p if k.startswith("enc.down") and k.endswith(".res0.norm1.w"))
        self._ws: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------------------------ primitive ops
    def _stream(self):
        return _lib.current_stream_ptr(self.device)

    def _conv(self, x, name, mode=EPI_BIAS, resid=None, stride=1, pad=None):
        cv = self.conv[name]
        B, H, W, Cin = x.shape
        assert Cin == cv.cin and x.is_contiguous(), (name, x.shape, cv.cin)
        if pad is None:
            pad = cv.k // 2
        Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
        out = torch.empty((B, Ho, Wo, cv.cout), dtype=torch.float32 if mode == EPI_F32 else torch.bfloat16, device=x.device)
        _lib.check(_lib.load().drag_conv2d_nhwc(_lib.ptr(x), B, H, W, Cin, _lib.ptr(cv.w), cv.cout, cv.k, stride, pad, Ho, Wo,
                                                mode, _lib.ptr(cv.b), _lib.ptr(out), _lib.ptr(resid), self._stream()),
                   "drag_conv2d_nhwc")
        return out

    def _gn(self, x, name, silu=True):
        B, H, W, Cc = x.shape
        need = B * 1024 * 2 * Cc + B * GN_GROUPS * 2
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.float32, device=x.device)
        g, b = self.norm[name]
        out = torch.empty_like(x)
        _lib.check(_lib.load().drag_groupnorm_nhwc(_lib.ptr(x), _lib.ptr(out), B, H * W, Cc, GN_GROUPS, _lib.ptr(g), _lib.ptr(b),
                                                   GN_EPS, int(silu), _lib.ptr(self._ws), self._ws.numel(), self._stream()),
                   "drag_groupnorm_nhwc")
        return out

    def _res(self, x, name):
        h = self._conv(self._gn(x, name + ".norm1"), name + ".conv1")
        h = self._gn(h, name + ".norm2")
        xs = self._conv(x, name + ".short") if (name + ".short") in self.conv else x
        return self._conv(h, name + ".conv2", mode=EPI_RESID, resid=xs)

    def _attn(self, x, name):
        B, H, W, Cc = x.shape
        N = H * W
        wq, bq, wk, bk, wv, wo, bo = self.lin[name]
        h = self._gn(x, name + ".norm", silu=False).view(B * N, Cc)
        out = torch.empty_like(x)
        rows_max = max(128, min(N, (_ATTN_SCORE_BYTES // (4 * N)) // 128 * 128))
        lib = _lib.load()
        Np = _pad_to(N, 32)                       # keys padded to the GEMM's N granularity; padded columns get P = 0
        for b in range(B):
            hb = h[b * N:(b + 1) * N]
            if Np != N:
                hp = torch.zeros((Np, Cc), dtype=torch.bfloat16, device=x.device)
                hp[:N] = hb
                hb = hp
            q = ops.linear(hb[:N], wq, bq)
            k = ops.linear(hb, wk, bk)                                 # [Np, C]
            vt = ops.linear(wv, hb)                                    # [C, Np] = Wv h^T (zero columns past N)
            xb, ob = x.view(B, N, Cc)[b], out.view(B, N, Cc)[b]
            for r0 in range(0, N, rows_max):
                r1 = min(N, r0 + rows_max)
                s = ops.linear(q[r0:r1], k, None, mode=ops.EPI_BIAS_F32)          # fp32 scores [rows, Np]
                pr = torch.zeros((r1 - r0, Np), dtype=torch.bfloat16, device=x.device)
                _lib.check(lib.drag_softmax_rows(_lib.ptr(s), Np, _lib.ptr(pr), Np, r1 - r0, N, self._stream()),
                           "drag_softmax_rows")
                o = ops.linear(pr, vt)
                ops.linear(o, wo, bo, mode=ops.EPI_GATE_RESID, resid=xb[r0:r1], out=ob[r0:r1])
        return out

    # ------------------------------------------------------------------------------------ public surface
    @torch.no_grad()
    def decode_nhwc(self, latents: torch.Tensor) -> torch.Tensor:
        """latents [B,16,h,w] (sampler space) -> decoder output fp32 NHWC [B,8h,8w,64] (channels 0..2 = RGB in ~[-1,1])."""
        lat = latents.to(self.device).contiguous()
        B, Cz, h, w = lat.shape
        lib = _lib.load()
        z = torch.empty((B, h, w, 64), dtype=torch.bfloat16, device=self.device)
        is_f32 = int(lat.dtype == torch.float32)
        if not is_f32:
            lat = lat.to(torch.bfloat16)
        _lib.check(lib.drag_nchw_to_nhwc_pad(_lib.ptr(lat), is_f32, _lib.ptr(z), B, Cz, h, w, 64, 1.0 / SCALE_FACTOR,
                                             SHIFT_FACTOR, self._stream()), "drag_nchw_to_nhwc_pad")
        x = self._conv(z, "dec.conv_in")
        x = self._res(x, "dec.mid.res0")
        x = self._attn(x, "dec.mid.attn")
        x = self._res(x, "dec.mid.res1")
        for L in range(self.n_up):
            for i in range(3):
                x = self._res(x, f"dec.up{L}.res{i}")
            if L != self.n_up - 1:
                B_, H_, W_, C_ = x.shape
                up = torch.empty((B_, 2 * H_, 2 * W_, C_), dtype=torch.bfloat16, device=self.device)
                _lib.check(lib.drag_upsample2x_nhwc(_lib.ptr(x), _lib.ptr(up), B_, H_, W_, C_, self._stream()),
                           "drag_upsample2x_nhwc")
                x = self._conv(up, f"dec.up{L}.upsample")
        return self._conv(self._gn(x, "dec.norm_out"), "dec.conv_out", mode=EPI_F32)

    @torch.no_grad()
    def decode(self, latents: torch.Tensor, output_type: str = "u8"):
        """"u8": uint8 [B,H,W,3] (what VaeImageProcessor hands to PIL); "pt": fp32 [B,3,H,W] decoder output; "pil": list."""
        y = self.decode_nhwc(latents)
        B, H, W, ld = y.shape
        lib = _lib.load()
        if output_type == "pt":
            out = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)
            _lib.check(lib.drag_nhwc_to_nchw_f32(_lib.ptr(y), 1, ld, _lib.ptr(out), B, 3, H, W, 1.0, 0.0, self._stream()),
                       "drag_nhwc_to_nchw_f32")
            return out
        u8 = torch.empty((B, H, W, 3), dtype=torch.uint8, device=self.device)
        _lib.check(lib.drag_image_postprocess_u8(_lib.ptr(y), ld, _lib.ptr(u8), B * H * W, self._stream()),
                   "drag_image_postprocess_u8")
        if output_type == "pil":
            from PIL import Image
            return [Image.fromarray(a) for a in u8.cpu().numpy()]
        return u8

    @torch.no_grad()
    def encode_moments(self, image: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """image uint8 [B,H,W,3] or float [B,3,H,W] in [-1,1] (H, W multiples of 8) -> moments fp32 [B,32,H/8,W/8].
        mask (uint8 [B,H,W], with a uint8 image only): non-zero pixels are zeroed first (FluxFill's masked image)."""
        lib = _lib.load()
        img = image.to(self.device).contiguous()
        if img.dtype == torch.uint8:
            B, H, W, _ = img.shape
            x = torch.empty((B, H, W, 64), dtype=torch.bfloat16, device=self.device)
            m = mask.to(self.device, torch.uint8).contiguous() if mask is not None else None
            _lib.check(lib.drag_image_preprocess_u8(_lib.ptr(img), _lib.ptr(m), _lib.ptr(x), B * H * W, 64, self._stream()),
                       "drag_image_preprocess_u8")
        else:
            B, _, H, W = img.shape
            img = img.float()
            x = torch.empty((B, H, W, 64), dtype=torch.bfloat16, device=self.device)
            _lib.check(lib.drag_nchw_to_nhwc_pad(_lib.ptr(img), 1, _lib.ptr(x), B, 3, H, W, 64, 1.0, 0.0, self._stream()),
                       "drag_nchw_to_nhwc_pad")
        if H % 8 or W % 8:
            raise ValueError("encode: image height and width must be multiples of 8")
        x = self._conv(x, "enc.conv_in")
        for L in range(self.n_down):
            for i in range(2):
                x = self._res(x, f"enc.down{L}.res{i}")
            if L != self.n_down - 1:
                x = self._conv(x, f"enc.down{L}.downsample", stride=2, pad=0)     # pad (0,1,0,1): right/bottom = zero fill
        x = self._res(x, "enc.mid.res0")
        x = self._attn(x, "enc.mid.attn")
        x = self._res(x, "enc.mid.res1")
        m = self._conv(self._gn(x, "enc.norm_out"), "enc.conv_out", mode=EPI_F32)   # [B,h,w,64] fp32, 32 used
        B_, h, w, ld = m.shape
        out = torch.empty((B_, 32, h, w), dtype=torch.float32, device=self.device)
        _lib.check(lib.drag_nhwc_to_nchw_f32(_lib.ptr(m), 1, ld, _lib.ptr(out), B_, 32, h, w, 1.0, 0.0, self._stream()),
                   "drag_nhwc_to_nchw_f32")
        return out

    @torch.no_grad()
    def encode(self, image: torch.Tensor, generator: Optional[torch.Generator] = None,
               mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(sample - shift) * scaling like the Flux pipelines; generator=None -> distribution mode. The Gaussian
        noise is drawn with the caller's generator on ITS device in the VAE dtype, bf16 (diffusers randn_tensor semantics).
        The reparameterisation itself is a handful of torch elementwise ops on the 16-channel latent, once per image."""
        mom = self.encode_moments(image, mask)
        mean, logvar = mom[:, :16], mom[:, 16:]
        if generator is None:
            z = mean
        else:
            if isinstance(generator, (list, tuple)):      # one generator per batch element (diffusers randn_tensor)
                if len(generator) != mean.shape[0]:
                    raise ValueError("encode: need one generator per image")
                noise = torch.cat([torch.randn(mean[i:i + 1].shape, generator=g, device=g.device, dtype=torch.bfloat16)
                                   for i, g in enumerate(generator)])
            else:
                noise = torch.randn(mean.shape, generator=generator, device=generator.device, dtype=torch.bfloat16)
            z = mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise.to(self.device).float()
        return ((z - SHIFT_FACTOR) * SCALE_FACTOR).to(torch.bfloat16)


def from_diffusers_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """AutoencoderKL (diffusers) key names -> the flat names of oracle/vae.py. Levels keep execution order."""
    out = {}

    def cp(dst, src, reshape_1x1=False):
        w = sd[src + ".weight"]
        if reshape_1x1 and w.dim() == 2:
            w = w[:, :, None, None]
        out[dst + ".w"], out[dst + ".b"] = w, sd[src + ".bias"]

    def res(dst, src):
        cp(dst + ".norm1", src + ".norm1"); cp(dst + ".conv1", src + ".conv1")
        cp(dst + ".norm2", src + ".norm2"); cp(dst + ".conv2", src + ".conv2")
        if src + ".conv_shortcut.weight" in sd:
            cp(dst + ".short", src + ".conv_shortcut")

    for side, mod in (("enc", "encoder"), ("dec", "decoder")):
        cp(f"{side}.conv_in", f"{mod}.conv_in"); cp(f"{side}.conv_out", f"{mod}.conv_out")
        cp(f"{side}.norm_out", f"{mod}.conv_norm_out")
        res(f"{side}.mid.res0", f"{mod}.mid_block.resnets.0"); res(f"{side}.mid.res1", f"{mod}.mid_block.resnets.1")
        a = f"{mod}.mid_block.attentions.0"
        cp(f"{side}.mid.attn.norm", a + ".group_norm")
        for d, s in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj", "to_out.0")):
            cp(f"{side}.mid.attn.{d}", f"{a}.{s}", reshape_1x1=True)
    L = 0
    while f"encoder.down_blocks.{L}.resnets.0.norm1.weight" in sd:
        for i in range(2):
            res(f"enc.down{L}.res{i}", f"encoder.down_blocks.{L}.resnets.{i}")
        if f"encoder.down_blocks.{L}.downsamplers.0.conv.weight" in sd:
            cp(f"enc.down{L}.downsample", f"encoder.down_blocks.{L}.downsamplers.0.conv")
        L += 1
    L = 0
    while f"decoder.up_blocks.{L}.resnets.0.norm1.weight" in sd:
        for i in range(3):
            res(f"dec.up{L}.res{i}", f"decoder.up_blocks.{L}.resnets.{i}")
        if f"decoder.up_blocks.{L}.upsamplers.0.conv.weight" in sd:
            cp(f"dec.up{L}.upsample", f"decoder.up_blocks.{L}.upsamplers.0.conv")
        L += 1
    return out
