"""GPU numerics: tcgen05 GEMM core + fused epilogues vs a plain PyTorch fp32 reference of the same op
(bf16 inputs, fp32 accumulate, bf16 output => tolerance = bf16 rounding of the result)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(lib):
    from domain_rag_b200 import ops
    return ops


def rnd(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).bfloat16()


def check(got, want, tol=2e-2):
    got, want = got.float(), want.float()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    rel = (got - want).norm().item() / max(want.norm().item(), 1e-12)
    assert err <= tol * max(ref, 1.0) and rel <= 1e-2, f"max abs err {err} (ref max {ref}), rel-L2 {rel}"


@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64), (128, 256, 128), (256, 512, 512), (100, 256, 320), (5337, 3072, 3072),
    (257, 768, 1024), (50, 2304, 768), (8, 3072, 256), (1000, 128, 192), (300, 64, 384), (77, 96, 72),
    (4096, 12288, 3072), (2265, 3072, 15360),
])
def test_plain_linear(ops, M, N, K):
    a, w, b = rnd((M, K), 1), rnd((N, K), 2, K ** -0.5), rnd((N,), 3)
    got = ops.linear(a, w, b)
    want = a.float() @ w.float().t() + b.float()
    check(got, want)
    got2 = ops.linear(a, w, None, mode=ops.EPI_BIAS_F32)
    check(got2, a.float() @ w.float().t(), tol=1e-2)


def test_strided_views(ops):
    big_a, big_w = rnd((300, 1024), 4), rnd((512, 2048), 5, 0.03)
    a, w = big_a[:, 256:768], big_w[:, 512:1024]          # row-strided, 16-byte aligned views
    check(ops.linear(a, w), a.float() @ w.float().t())
    out = torch.zeros((300, 1024), device="cuda", dtype=torch.bfloat16)
    ops.linear(a, w, out=out[:, 512:])                     # write into a column slice (concat-free)
    check(out[:, 512:], a.float() @ w.float().t())
    assert float(out[:, :512].abs().max()) == 0.0


@pytest.mark.parametrize("mode", ["gelu_tanh", "quick_gelu", "silu"])
def test_activations(ops, mode):
    a, w, b = rnd((333, 512), 6), rnd((768, 512), 7, 512 ** -0.5), rnd((768,), 8)
    y = a.float() @ w.float().t() + b.float()
    if mode == "gelu_tanh":
        want, m = torch.nn.functional.gelu(y, approximate="tanh"), ops.EPI_GELU_TANH
    elif mode == "quick_gelu":
        want, m = y * torch.sigmoid(1.702 * y), ops.EPI_QUICK_GELU
    else:
        want, m = torch.nn.functional.silu(y), ops.EPI_SILU
    check(ops.linear(a, w, b, mode=m), want)


def test_gated_residual(ops):
    B, S, d = 3, 200, 512
    a, w, b = rnd((B * S, 768), 9), rnd((d, 768), 10, 768 ** -0.5), rnd((d,), 11)
    x, gate = rnd((B * S, d), 12), rnd((B, d), 13)
    want = x.float() + gate.float().repeat_interleave(S, 0) * (a.float() @ w.float().t() + b.float())
    got = ops.linear(a, w, b, mode=ops.EPI_GATE_RESID, resid=x, gate=gate, rows_per_batch=S, out=x.clone())
    check(got, want, tol=3e-2)
    got2 = ops.linear(a, w, b, mode=ops.EPI_GATE_RESID, resid=x)   # plain residual (ViT blocks)
    check(got2, x.float() + a.float() @ w.float().t() + b.float(), tol=3e-2)


def test_qkv_rmsnorm_rope(ops):
    B, S_txt, S_img, H, K = 2, 77, 130, 4, 512
    S = S_txt + S_img
    a_img, a_txt = rnd((B * S_img, K), 14), rnd((B * S_txt, K), 15)
    w, bias = rnd((3 * H * 128, K), 16, K ** -0.5), rnd((3 * H * 128,), 17, 0.1)
    qw, kw = (1 + 0.1 * rnd((128,), 18).float()).bfloat16(), (1 + 0.1 * rnd((128,), 19).float()).bfloat16()
    g = torch.Generator(device="cuda").manual_seed(20)
    ang = torch.rand((S, 64), generator=g, device="cuda") * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    q = torch.zeros((B, H, S, 128), device="cuda", dtype=torch.bfloat16)
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    ops.qkv_rope(a_txt, w, bias, q, k, v, qw, kw, cos, sin, 0, S_txt)
    ops.qkv_rope(a_img, w, bias, q, k, v, qw, kw, cos, sin, S_txt, S_img)

    def ref(a, rows, off):
        y = (a.float() @ w.float().t() + bias.float()).bfloat16().float().view(B, rows, 3, H, 128)
        outs = []
        for i, nw in ((0, qw), (1, kw)):
            x = y[:, :, i]
            x = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().float() * nw.float()
            x = x.bfloat16().float()
            c, s = cos[off:off + rows][None, :, None, :], sin[off:off + rows][None, :, None, :]
            x0, x1 = x[..., 0::2], x[..., 1::2]
            outs.append(torch.stack([x0 * c - x1 * s, x1 * c + x0 * s], -1).flatten(-2))
        outs.append(y[:, :, 2])
        return [o.permute(0, 2, 1, 3) for o in outs]   # [B,H,rows,128]

    for off, rows, a in ((0, S_txt, a_txt), (S_txt, S_img, a_img)):
        rq, rk, rv = ref(a, rows, off)
        check(q[:, :, off:off + rows], rq)
        check(k[:, :, off:off + rows], rk)
        check(v[:, :, off:off + rows], rv)
