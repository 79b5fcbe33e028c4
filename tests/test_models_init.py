"""CPU: the product's seeded random-init helpers produce exactly the oracle's parameter sets (so GPU parity tests and
the CLIs build the same weights from a seed without the product importing oracle/)."""
import torch

from domain_rag_b200 import siglip as S
from domain_rag_b200 import vae_init as VI
from oracle import siglip as OS
from oracle import vae as OV


def test_vae_init_equals_oracle():
    a, b = VI.init_params(seed=5000, ch=32), OV.init_params(seed=5000, ch=32)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert VI.param_shapes() == OV.param_shapes()


def test_siglip_redux_init_equals_oracle():
    cfgd = dict(hidden=144, layers=2, heads=2, mlp=272, patch=14, image=56)
    a, b = VI.init_siglip(S.SiglipConfig(**cfgd), 6000), OS.init_state(OS.SiglipConfig(**cfgd), 6000)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    r1, r2 = VI.init_redux(6100, 144, 192, 64), OS.init_redux(6100, 144, 192, 64)
    assert all(torch.equal(r1[k], r2[k]) for k in r1)
