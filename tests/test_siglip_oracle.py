"""CPU: the SigLIP oracle against transformers.SiglipVisionModel (independent code for the same architecture)."""
import pytest
import torch

from oracle import siglip as OS


@pytest.mark.parametrize("cfg", [OS.SiglipConfig(hidden=144, layers=2, heads=2, mlp=272, patch=14, image=56),
                                 OS.SiglipConfig(hidden=128, layers=3, heads=4, mlp=200, patch=16, image=64)])
def test_oracle_matches_transformers_siglip(cfg):
    from transformers import SiglipVisionConfig, SiglipVisionModel
    hf = SiglipVisionConfig(hidden_size=cfg.hidden, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers,
                            num_attention_heads=cfg.heads, image_size=cfg.image, patch_size=cfg.patch,
                            hidden_act="gelu_pytorch_tanh", layer_norm_eps=cfg.eps)
    m = SiglipVisionModel(hf).eval()
    state = OS.init_state(cfg, seed=3)
    missing, unexpected = m.load_state_dict(state, strict=False)
    assert not unexpected, unexpected
    assert all(("head." in k) or ("position_ids" in k) for k in missing), missing       # pooling head is unused by Redux
    x = torch.randn(2, 3, cfg.image, cfg.image, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = m(pixel_values=x).last_hidden_state
        got = OS.last_hidden_state(state, cfg, x)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def test_full_size_inventory_and_redux():
    cfg = OS.SO400M
    assert cfg.tokens == 729 and cfg.hidden // cfg.heads == 72
    r = OS.init_redux(d_in=64, d_hidden=192, d_out=32)
    t = torch.randn(2, 5, 64, generator=torch.Generator().manual_seed(0))
    out = OS.redux_embed(r, t)
    assert out.shape == (2, 5, 32)
    want = torch.nn.functional.silu(t @ r["redux_up.weight"].t() + r["redux_up.bias"]) @ r["redux_down.weight"].t() + r["redux_down.bias"]
    torch.testing.assert_close(out, want)
