"""GPU parity: SigLIP vision tower (head dim 72 padded to 128, MLP 4304 padded) and the Redux embedder vs the
CPU fp32 oracle on the same bf16-rounded weights."""
import pytest
import torch

from oracle import siglip as OS

pytestmark = pytest.mark.gpu


def rel_l2(got, want):
    return ((got.float() - want.float()).norm() / want.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("cfgd,B", [(dict(hidden=160, layers=2, heads=2, mlp=272, patch=14, image=60), 3),
                                    (dict(hidden=1152, layers=2, heads=16, mlp=4304, patch=14, image=384), 1)])
def test_siglip_tower_matches_oracle(lib, cfgd, B):
    from domain_rag_b200 import siglip as S
    ocfg, cfg = OS.SiglipConfig(**cfgd), S.SiglipConfig(**cfgd)
    state = {k: v.bfloat16().float() for k, v in OS.init_state(ocfg, seed=6000).items()}
    x = torch.randn(B, 3, cfg.image, cfg.image, generator=torch.Generator().manual_seed(2)).clamp(-1, 1)
    want = OS.last_hidden_state(state, ocfg, x)
    got = S.SiglipVisionTower(cfg, state).last_hidden_state(x.cuda()).float().cpu()
    assert got.shape == want.shape
    assert rel_l2(got, want) < 2e-2, rel_l2(got, want)


def test_redux_embedder_matches_oracle(lib):
    from domain_rag_b200 import siglip as S
    r = {k: v.bfloat16().float() for k, v in OS.init_redux(seed=6100, d_in=1152, d_hidden=1536, d_out=512).items()}
    t = torch.randn(2, 729, 1152, generator=torch.Generator().manual_seed(4)).bfloat16()
    want = OS.redux_embed(r, t.float())
    got = S.ReduxImageEncoder(r)(t.cuda()).float().cpu()
    assert rel_l2(got, want) < 1e-2


def test_preprocess_matches_oracle():
    import numpy as np
    from PIL import Image
    from domain_rag_b200 import siglip as S
    im = Image.fromarray((np.random.default_rng(0).random((100, 150, 3)) * 255).astype(np.uint8))
    assert torch.equal(S.preprocess([im]), OS.preprocess([im]))
