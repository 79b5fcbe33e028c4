"""CPU: diffusers FluxTransformer2DModel key layout -> fused engine layout. The synthetic checkpoint is built by
splitting a seeded parameter set back into diffusers' per-projection tensors; the conversion must reproduce the
original set exactly (and reject shape mismatches)."""
import pytest
import torch

from domain_rag_b200 import flux as F
from oracle import flux as OF


def to_diffusers(p, cfg):
    d = cfg.d
    sd = {}

    def lin(dst, w, b):
        sd[dst + ".weight"], sd[dst + ".bias"] = p[w], p[b]

    lin("x_embedder", "x_in.w", "x_in.b"); lin("context_embedder", "ctx_in.w", "ctx_in.b"); lin("proj_out", "final.w", "final.b")
    for a, b in (("t_in", "timestep_embedder"), ("g_in", "guidance_embedder"), ("p_in", "text_embedder")):
        lin(f"time_text_embed.{b}.linear_1", a + ".w1", a + ".b1"); lin(f"time_text_embed.{b}.linear_2", a + ".w2", a + ".b2")
    off = 0

    def take(n):
        nonlocal off
        w, b = p["mod.w"][off:off + n], p["mod.b"][off:off + n]
        off += n
        return w, b

    for i in range(cfg.n_double):
        blk = f"transformer_blocks.{i}."
        for st, norm, names, o, nq, nk, ff in (("img", "norm1", ("attn.to_q", "attn.to_k", "attn.to_v"), "attn.to_out.0", "attn.norm_q", "attn.norm_k", "ff"),
                                               ("txt", "norm1_context", ("attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj"), "attn.to_add_out",
                                                "attn.norm_added_q", "attn.norm_added_k", "ff_context")):
            q = f"double.{i}.{st}."
            sd[blk + norm + ".linear.weight"], sd[blk + norm + ".linear.bias"] = take(6 * d)
            for j, n in enumerate(names):
                sd[blk + n + ".weight"], sd[blk + n + ".bias"] = p[q + "qkv.w"][j * d:(j + 1) * d], p[q + "qkv.b"][j * d:(j + 1) * d]
            sd[blk + nq + ".weight"], sd[blk + nk + ".weight"] = p[q + "qnorm"], p[q + "knorm"]
            lin(blk + o, q + "out.w", q + "out.b"); lin(blk + ff + ".net.0.proj", q + "mlp1.w", q + "mlp1.b"); lin(blk + ff + ".net.2", q + "mlp2.w", q + "mlp2.b")
    for i in range(cfg.n_single):
        blk, q = f"single_transformer_blocks.{i}.", f"single.{i}."
        sd[blk + "norm.linear.weight"], sd[blk + "norm.linear.bias"] = take(3 * d)
        for j, n in enumerate("qkv"):
            sd[blk + f"attn.to_{n}.weight"], sd[blk + f"attn.to_{n}.bias"] = p[q + "qkv.w"][j * d:(j + 1) * d], p[q + "qkv.b"][j * d:(j + 1) * d]
        sd[blk + "attn.norm_q.weight"], sd[blk + "attn.norm_k.weight"] = p[q + "qnorm"], p[q + "knorm"]
        lin(blk + "proj_mlp", q + "mlp.w", q + "mlp.b"); lin(blk + "proj_out", q + "out.w", q + "out.b")
    sd["norm_out.linear.weight"], sd["norm_out.linear.bias"] = take(2 * d)
    return sd


def test_roundtrip_and_shape_check():
    small = dict(in_channels=384, d=256, heads=2, n_double=2, n_single=3, txt_dim=64, pooled_dim=32, out_channels=64, guidance=True)
    p = OF.init_params(OF.FluxConfig(**small), seed=9)
    cfg = F.FluxConfig(**small)
    sd = to_diffusers(p, cfg)
    got = F.from_diffusers_state_dict(sd, cfg)
    assert set(got) == set(F.param_shapes(cfg))
    for k in got:
        assert torch.equal(got[k], p[k]), k
    sd["x_embedder.weight"] = sd["x_embedder.weight"][:, :64]
    with pytest.raises(ValueError):
        F.from_diffusers_state_dict(sd, cfg)
