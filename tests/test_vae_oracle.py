"""CPU: the VAE oracle against an independent implementation of the same architecture
(torchtitan.experiments.flux.model.autoencoder, BFL layout) with a weight remap."""
import pytest
import torch

from oracle import vae as OV


def _to_bfl(p, ae):
    """Copy oracle parameters into torchtitan's AutoEncoder (BFL naming; decoder `up` list is reversed)."""
    sd = {}

    def conv(dst, src):
        sd[dst + ".weight"], sd[dst + ".bias"] = p[src + ".w"], p[src + ".b"]

    def res(dst, src):
        for n in ("norm1", "norm2"):
            sd[f"{dst}.{n}.weight"], sd[f"{dst}.{n}.bias"] = p[f"{src}.{n}.w"], p[f"{src}.{n}.b"]
        conv(dst + ".conv1", src + ".conv1")
        conv(dst + ".conv2", src + ".conv2")
        if src + ".short.w" in p:
            conv(dst + ".nin_shortcut", src + ".short")

    for side, mod in (("enc", "encoder"), ("dec", "decoder")):
        conv(f"{mod}.conv_in", f"{side}.conv_in")
        conv(f"{mod}.conv_out", f"{side}.conv_out")
        sd[f"{mod}.norm_out.weight"], sd[f"{mod}.norm_out.bias"] = p[f"{side}.norm_out.w"], p[f"{side}.norm_out.b"]
        res(f"{mod}.mid.block_1", f"{side}.mid.res0")
        res(f"{mod}.mid.block_2", f"{side}.mid.res1")
        sd[f"{mod}.mid.attn_1.norm.weight"], sd[f"{mod}.mid.attn_1.norm.bias"] = p[f"{side}.mid.attn.norm.w"], p[f"{side}.mid.attn.norm.b"]
        for a, b in (("q", "q"), ("k", "k"), ("v", "v"), ("proj_out", "proj")):
            conv(f"{mod}.mid.attn_1.{a}", f"{side}.mid.attn.{b}")
    for L in range(4):
        for i in range(2):
            res(f"encoder.down.{L}.block.{i}", f"enc.down{L}.res{i}")
        if L != 3:
            conv(f"encoder.down.{L}.downsample.conv", f"enc.down{L}.downsample")
        for i in range(3):
            res(f"decoder.up.{3 - L}.block.{i}", f"dec.up{L}.res{i}")
        if L != 3:
            conv(f"decoder.up.{3 - L}.upsample.conv", f"dec.up{L}.upsample")
    missing, unexpected = ae.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


@pytest.mark.parametrize("ch", [32])
def test_oracle_matches_torchtitan_autoencoder(ch):
    A = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    p = OV.init_params(seed=5000, ch=ch)
    ae = A.AutoEncoder(A.AutoEncoderParams(resolution=64, ch=ch)).float().eval()
    _to_bfl(p, ae)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(2, 16, 8, 12, generator=g)
    x = torch.rand(1, 3, 48, 64, generator=g) * 2 - 1
    with torch.no_grad():
        torch.testing.assert_close(OV.decode_latents(z, p), ae.decode(z), rtol=1e-4, atol=1e-4)
        ae.reg.sample = False
        torch.testing.assert_close(OV.encode_image(x, p), ae.encode(x), rtol=1e-4, atol=1e-4)
        mom = OV.encoder(x, p)
        assert mom.shape == (1, 32, 6, 8)
        img = OV.decode_latents(z, p)
        u8 = OV.postprocess_u8(img)
        assert u8.shape == (2, 64, 96, 3) and u8.dtype == torch.uint8
        back = OV.preprocess_image(u8)
        assert back.shape == (2, 3, 64, 96) and float(back.min()) >= -1 and float(back.max()) <= 1


def test_param_inventory_full_size():
    s = OV.param_shapes()
    n = sum(int(torch.tensor(v).prod()) for v in s.values())
    assert 83_000_000 < n < 84_500_000          # FLUX.1 VAE: 83.8 M parameters
    assert s["dec.conv_in.w"] == (512, 16, 3, 3) and s["dec.conv_out.w"] == (3, 128, 3, 3)
    assert s["dec.up2.res0.short.w"] == (256, 512, 1, 1) and s["enc.conv_out.w"] == (32, 512, 3, 3)
