"""GPU parity: Flux VAE decode / encode (implicit-GEMM convolutions on the tcgen05 core + GroupNorm / upsample /
softmax kernels) vs the CPU fp32 oracle on the same bf16-rounded weights, plus the individual kernels vs PyTorch."""
import pytest
import torch
import torch.nn.functional as F

from oracle import vae as OV

pytestmark = pytest.mark.gpu


def rel_l2(got, want):
    return ((got.float() - want.float()).norm() / want.float().norm().clamp_min(1e-12)).item()


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).bfloat16()


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride", [
    (1, 16, 128, 64, 128, 3, 1),      # one 128-pixel row segment per tile
    (2, 8, 64, 128, 256, 3, 1),       # two full rows per tile, CTA-pair kernel
    (1, 24, 200, 64, 64, 3, 1),       # ragged row (200 = 128 + 72), single-CTA kernel (C_out 64)
    (1, 12, 16, 128, 128, 1, 1),      # 1x1 shortcut, 8 rows per tile, H not a multiple of rows-per-tile
    (2, 32, 48, 64, 128, 3, 2),       # stride-2 downsample with (0,1,0,1) padding, W_out = 24 -> ragged tile
    (1, 64, 256, 256, 256, 3, 1),     # several tiles per row, BN = 256 pair tiles
])
def test_conv2d_matches_torch(lib, B, H, W, Cin, Cout, k, stride):
    from domain_rag_b200 import _lib
    x = rnd((B, H, W, Cin), 1)
    w = rnd((Cout, Cin, k, k), 2, (Cin * k * k) ** -0.5)
    bias = rnd((Cout,), 3)
    xt = x.float().permute(0, 3, 1, 2)
    if stride == 2:
        want = F.conv2d(F.pad(xt, (0, 1, 0, 1)), w.float(), bias.float(), stride=2)
        pad, Ho, Wo = 0, H // 2, W // 2
    else:
        want = F.conv2d(xt, w.float(), bias.float(), padding=k // 2)
        pad, Ho, Wo = k // 2, H, W
    want = want.permute(0, 2, 3, 1)
    wt = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous().cuda()
    out = torch.full((B, Ho, Wo, Cout), 7.0, dtype=torch.bfloat16, device="cuda")
    xc, bc = x.cuda(), bias.cuda()
    _lib.check(lib.drag_conv2d_nhwc(_lib.ptr(xc), B, H, W, Cin, _lib.ptr(wt), Cout, k, stride, pad, Ho, Wo, 0,
                                    _lib.ptr(bc), _lib.ptr(out), None, _lib.current_stream_ptr(xc.device)), "conv")
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), want) < 1e-2
    assert (out.cpu().float() - want).abs().max().item() < 3e-2 * max(1.0, want.abs().max().item())
    # residual epilogue
    res = rnd((B, Ho, Wo, Cout), 4).cuda()
    out2 = torch.empty_like(out)
    _lib.check(lib.drag_conv2d_nhwc(_lib.ptr(xc), B, H, W, Cin, _lib.ptr(wt), Cout, k, stride, pad, Ho, Wo, 4,
                                    _lib.ptr(bc), _lib.ptr(out2), _lib.ptr(res), _lib.current_stream_ptr(xc.device)), "conv")
    torch.cuda.synchronize()
    assert rel_l2(out2.cpu(), want + res.cpu().float()) < 1e-2


@pytest.mark.parametrize("B,HW,C,silu", [(2, 96, 128, True), (1, 4096, 512, True), (3, 1000, 256, False), (1, 7, 64, True)])
def test_groupnorm_silu_matches_torch(lib, B, HW, C, silu):
    from domain_rag_b200 import _lib
    x = (rnd((B, HW, C), 5, 2.0).float() + 0.7).bfloat16()
    g, b = (1 + 0.1 * rnd((C,), 6).float()).bfloat16(), rnd((C,), 7, 0.1)
    want = F.group_norm(x.float().permute(0, 2, 1), 32, g.float(), b.float(), 1e-6)
    if silu:
        want = F.silu(want)
    want = want.permute(0, 2, 1)
    ws = torch.empty(B * 1024 * 2 * C + B * 64, dtype=torch.float32, device="cuda")
    out = torch.empty((B, HW, C), dtype=torch.bfloat16, device="cuda")
    xc = x.cuda()
    gc, bc = g.cuda(), b.cuda()          # keep the device copies alive for the launch
    _lib.check(lib.drag_groupnorm_nhwc(_lib.ptr(xc), _lib.ptr(out), B, HW, C, 32, _lib.ptr(gc), _lib.ptr(bc), 1e-6,
                                       int(silu), _lib.ptr(ws), ws.numel(), _lib.current_stream_ptr(xc.device)), "gn")
    torch.cuda.synchronize()
    assert (out.cpu().float() - want).abs().max().item() < 3e-2
    assert rel_l2(out.cpu(), want) < 5e-3


def test_upsample_softmax_layout_pixels(lib):
    from domain_rag_b200 import _lib
    st = _lib.current_stream_ptr(torch.device("cuda", 0))
    x = rnd((2, 5, 7, 64), 8).cuda()
    up = torch.empty((2, 10, 14, 64), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.drag_upsample2x_nhwc(_lib.ptr(x), _lib.ptr(up), 2, 5, 7, 64, st), "up")
    want = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), want)
    s = (torch.randn(37, 96, generator=torch.Generator().manual_seed(9)) * 4).cuda()
    p = torch.zeros((37, 96), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.drag_softmax_rows(_lib.ptr(s), 96, _lib.ptr(p), 96, 37, 92, st), "softmax")
    assert float(p[:, 92:].abs().max()) == 0.0
    assert (p[:, :92].float() - torch.softmax(s[:, :92], -1)).abs().max().item() < 4e-3
    z = rnd((2, 16, 6, 10), 10)
    nh = torch.empty((2, 6, 10, 64), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.drag_nchw_to_nhwc_pad(_lib.ptr(z.cuda()), 0, _lib.ptr(nh), 2, 16, 6, 10, 64, 2.0, 0.5, st), "layout")
    assert float(nh[..., 16:].abs().max()) == 0.0
    assert (nh[..., :16].float().cpu() - (z.float() * 2.0 + 0.5).permute(0, 2, 3, 1)).abs().max().item() < 2e-2
    img = torch.randn(1, 9, 11, 8, generator=torch.Generator().manual_seed(11)).cuda()
    u8 = torch.empty((1, 9, 11, 3), dtype=torch.uint8, device="cuda")
    _lib.check(lib.drag_image_postprocess_u8(_lib.ptr(img), 8, _lib.ptr(u8), 99, st), "post")
    want = ((img[..., :3] / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
    assert torch.equal(u8, want)
    back = torch.empty((1, 9, 11, 64), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.drag_image_preprocess_u8(_lib.ptr(u8), None, _lib.ptr(back), 99, 64, st), "pre")
    assert (back[..., :3].float() - (u8.float() / 255 * 2 - 1)).abs().max().item() < 4e-3 and float(back[..., 3:].abs().max()) == 0


@pytest.mark.parametrize("ch,h,w", [(64, 8, 12), (64, 16, 16)])
def test_vae_decode_encode_match_oracle(lib, ch, h, w):
    from domain_rag_b200.vae import FluxVAE
    p = {k: v.bfloat16().float() for k, v in OV.init_params(seed=5000, ch=ch).items()}
    vae = FluxVAE(p, "cuda")
    g = torch.Generator().manual_seed(1)
    z = torch.randn(2, 16, h, w, generator=g).bfloat16()
    want = OV.decode_latents(z.float(), p)
    got = vae.decode(z.cuda(), output_type="pt").cpu()
    assert got.shape == want.shape
    assert rel_l2(got, want) < 3e-2, rel_l2(got, want)
    u8 = vae.decode(z.cuda()).cpu()
    want_u8 = OV.postprocess_u8(want)
    assert u8.shape == want_u8.shape
    assert (u8.int() - want_u8.int()).abs().float().mean().item() < 2.0     # bf16 activations: ~1 grey level
    img = torch.rand(1, 3, 8 * h, 8 * w, generator=g) * 2 - 1
    want_m = OV.encoder(img, p)
    got_m = vae.encode_moments(img.cuda()).cpu()
    assert rel_l2(got_m, want_m) < 3e-2, rel_l2(got_m, want_m)
    lat = vae.encode(img.cuda()).float().cpu()
    assert rel_l2(lat, OV.encode_image(img, p)) < 3e-2
    gen = torch.Generator().manual_seed(3)
    lat_s = vae.encode(img.cuda(), generator=gen).float().cpu()
    noise = torch.randn(want_m[:, :16].shape, generator=torch.Generator().manual_seed(3), dtype=torch.bfloat16).float()
    assert rel_l2(lat_s, OV.encode_image(img, p, noise=noise)) < 3e-2
