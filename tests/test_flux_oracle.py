"""CPU: the Flux oracle against the independent BFL-style implementation shipped with torchtitan
(weight remap: fused qkv / linear1 / linear2, last-layer (shift, scale) order), plus invariants of
the sampler glue. The reference's own diffusers code is not available offline (parity unpinned)."""
import pytest
import torch

from oracle import flux as OF

tt = pytest.importorskip("torchtitan.experiments.flux.model.model")
from torchtitan.experiments.flux.model.args import FluxModelArgs  # noqa: E402


def build_pair(seed=0):
    cfg = OF.FluxConfig(in_channels=64, d=256, heads=2, n_double=2, n_single=2, txt_dim=48, pooled_dim=32,
                        out_channels=64, guidance=False)
    p = OF.init_params(cfg, seed=seed)
    args = FluxModelArgs(in_channels=64, out_channels=64, vec_in_dim=32, context_in_dim=48, hidden_size=256,
                         mlp_ratio=4.0, num_heads=2, depth=2, depth_single_blocks=2, axes_dim=(16, 56, 56),
                         theta=10000, qkv_bias=True)
    m = tt.FluxModel(args).float().eval()
    d = cfg.d
    sd = {}
    sd["img_in.weight"], sd["img_in.bias"] = p["x_in.w"], p["x_in.b"]
    sd["txt_in.weight"], sd["txt_in.bias"] = p["ctx_in.w"], p["ctx_in.b"]
    for a, b in (("time_in", "t_in"), ("vector_in", "p_in")):
        sd[f"{a}.in_layer.weight"], sd[f"{a}.in_layer.bias"] = p[f"{b}.w1"], p[f"{b}.b1"]
        sd[f"{a}.out_layer.weight"], sd[f"{a}.out_layer.bias"] = p[f"{b}.w2"], p[f"{b}.b2"]
    for i in range(cfg.n_double):
        for st in ("img", "txt"):
            o = cfg.mod_offset_double(i, st == "txt")
            q = f"double.{i}.{st}."
            t = f"double_blocks.{i}.{st}_"
            sd[t + "mod.lin.weight"], sd[t + "mod.lin.bias"] = p["mod.w"][o:o + 6 * d], p["mod.b"][o:o + 6 * d]
            sd[t + "attn.qkv.weight"], sd[t + "attn.qkv.bias"] = p[q + "qkv.w"], p[q + "qkv.b"]
            sd[t + "attn.norm.query_norm.weight"], sd[t + "attn.norm.key_norm.weight"] = p[q + "qnorm"], p[q + "knorm"]
            sd[t + "attn.proj.weight"], sd[t + "attn.proj.bias"] = p[q + "out.w"], p[q + "out.b"]
            sd[t + "mlp.0.weight"], sd[t + "mlp.0.bias"] = p[q + "mlp1.w"], p[q + "mlp1.b"]
            sd[t + "mlp.2.weight"], sd[t + "mlp.2.bias"] = p[q + "mlp2.w"], p[q + "mlp2.b"]
    for i in range(cfg.n_single):
        o = cfg.mod_offset_single(i)
        q, t = f"single.{i}.", f"single_blocks.{i}."
        sd[t + "modulation.lin.weight"], sd[t + "modulation.lin.bias"] = p["mod.w"][o:o + 3 * d], p["mod.b"][o:o + 3 * d]
        sd[t + "linear1.weight"] = torch.cat([p[q + "qkv.w"], p[q + "mlp.w"]], 0)
        sd[t + "linear1.bias"] = torch.cat([p[q + "qkv.b"], p[q + "mlp.b"]], 0)
        sd[t + "linear2.weight"], sd[t + "linear2.bias"] = p[q + "out.w"], p[q + "out.b"]
        sd[t + "norm.query_norm.weight"], sd[t + "norm.key_norm.weight"] = p[q + "qnorm"], p[q + "knorm"]
    o = cfg.mod_offset_final()
    scale_w, shift_w = p["mod.w"][o:o + d], p["mod.w"][o + d:o + 2 * d]
    scale_b, shift_b = p["mod.b"][o:o + d], p["mod.b"][o + d:o + 2 * d]
    sd["final_layer.adaLN_modulation.1.weight"] = torch.cat([shift_w, scale_w], 0)   # BFL order: (shift, scale)
    sd["final_layer.adaLN_modulation.1.bias"] = torch.cat([shift_b, scale_b], 0)
    sd["final_layer.linear.weight"], sd["final_layer.linear.bias"] = p["final.w"], p["final.b"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("pe_embedder" in k or "norm" in k for k in missing), (missing, unexpected)
    return cfg, p, m


def test_oracle_matches_torchtitan_flux():
    cfg, p, m = build_pair()
    g = torch.Generator().manual_seed(5)
    B, h2, w2, s_txt = 2, 6, 8, 20
    x = torch.randn(B, h2 * w2, 64, generator=g)
    ctx = torch.randn(B, s_txt, 48, generator=g)
    pooled = torch.randn(B, 32, generator=g)
    t = torch.tensor([0.9, 0.3])
    img_ids, txt_ids = OF.image_ids(h2, w2), torch.zeros(s_txt, 3)
    with torch.no_grad():
        want = m(img=x, img_ids=img_ids[None].expand(B, -1, -1), txt=ctx, txt_ids=txt_ids[None].expand(B, -1, -1),
                 timesteps=t, y=pooled)
        got = OF.flux_forward(p, cfg, x, ctx, pooled, t, None, img_ids, txt_ids)
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)


def test_pack_unpack_and_schedule():
    z = torch.arange(2 * 16 * 8 * 12, dtype=torch.float32).view(2, 16, 8, 12)
    from einops import rearrange
    want = rearrange(z, "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)
    assert torch.equal(OF.pack_latents(z), want)
    assert torch.equal(OF.unpack_latents(OF.pack_latents(z), 8, 12), z)
    sig = OF.flow_match_sigmas(50, 4096)
    assert sig.shape == (51,) and sig[0] == pytest.approx(1.0) and sig[-1] == 0.0
    assert torch.all(sig[1:] < sig[:-1])
    # mu = 1.15 at 1024^2 (4096 tokens): sigma'(0.5) = e^mu / (e^mu + 1)
    import math
    mid = math.exp(1.15) / (math.exp(1.15) + (1 / 0.5 - 1))
    assert float(OF.flow_match_sigmas(2, 4096)[1]) == pytest.approx(mid, rel=1e-6)
    ids = OF.image_ids(3, 4)
    assert ids[:, 0].abs().sum() == 0 and ids[5].tolist() == [0.0, 1.0, 1.0]


def test_product_and_oracle_layouts_agree():
    from domain_rag_b200 import flux as F
    for kw in (dict(), dict(in_channels=384), dict(guidance=False, d=256, heads=2, n_double=1, n_single=1)):
        assert F.param_shapes(F.FluxConfig(**kw)) == OF.param_shapes(OF.FluxConfig(**kw))
        assert F.FluxConfig(**kw).n_mod == OF.FluxConfig(**kw).n_mod
    order = F.param_order(F.FluxConfig())
    assert len(order) == 20 + 20 * 19 + 8 * 38 and len(set(order)) == len(order)
    assert F.flow_match_sigmas(50, 4096) == pytest.approx(OF.flow_match_sigmas(50, 4096).tolist())


def test_sampler_glue_matches_torchtitan_sampling():
    """Second statement of the sampler glue (oracle/pipelines.py and the schedule had no independent check):
    torchtitan's BFL-style sampling.py - get_schedule (time_shift of linspace(1, 0, T+1), mu from the token count),
    position encodings, 2x2 latent packing and the Euler update `latents + (t_prev - t_curr) * pred` (:38-66, :195-209) -
    against oracle.flux.sample / flow_match_sigmas / image_ids / pack_latents on the same model and noise."""
    from torchtitan.experiments.flux import sampling as TS
    from torchtitan.experiments.flux import utils as TU
    cfg, p, m = build_pair(seed=3)
    T, H, W, s_txt = 5, 64, 96, 12
    h, w = H // 8, W // 8
    seq = (h // 2) * (w // 2)
    sched = TS.get_schedule(T, seq, shift=True)
    assert OF.flow_match_sigmas(T, seq).tolist() == pytest.approx(sched, rel=1e-6, abs=1e-7)
    assert torch.equal(TU.create_position_encoding_for_latents(1, h, w)[0], OF.image_ids(h // 2, w // 2))
    g = torch.Generator().manual_seed(9)
    z = torch.randn(2, 16, h, w, generator=g)
    ctx, pooled = torch.randn(2, s_txt, 48, generator=g), torch.randn(2, 32, generator=g)
    assert torch.equal(TU.pack_latents(z), OF.pack_latents(z))
    lat = TU.pack_latents(z)
    pos = TU.create_position_encoding_for_latents(2, h, w)
    with torch.no_grad():
        for t_curr, t_prev in zip(sched[:-1], sched[1:]):
            pred = m(img=lat, img_ids=pos, txt=ctx, txt_ids=torch.zeros(2, s_txt, 3), y=pooled,
                     timesteps=torch.full((2,), t_curr))
            lat = lat + (t_prev - t_curr) * pred
        want = TU.unpack_latents(lat, h, w)
        got = OF.unpack_latents(OF.sample(p, cfg, OF.pack_latents(z), ctx, pooled, 0.0, T, h // 2, w // 2), h, w)
    torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-3)


def test_guidance_embedder_is_the_timestep_embedder_applied_to_the_guidance_scale():
    """The guidance branch has no second implementation offline (torchtitan's FluxModel is the schnell variant). What can be
    pinned: diffusers' CombinedTimestepGuidanceTextProjEmbeddings adds THREE terms - timestep MLP(sinusoid(1000 t)), guidance
    MLP(sinusoid(1000 g)) with the same sinusoid, pooled-text MLP - so with g_in := t_in and g := t the conditioning vector must
    equal 2 * t_emb + p_emb, and a model without the branch must equal the same model with a zeroed guidance MLP."""
    cfg = OF.FluxConfig(in_channels=64, d=256, heads=2, n_double=1, n_single=1, txt_dim=48, pooled_dim=32, guidance=True)
    p = OF.init_params(cfg, seed=4)
    for k in ("w1", "b1", "w2", "b2"):
        p[f"g_in.{k}"] = p[f"t_in.{k}"].clone()
    t = torch.tensor([0.7, 0.2])
    pooled = torch.randn(2, 32, generator=torch.Generator().manual_seed(1))
    vec = OF.temb_vector(p, cfg, t, t, pooled)
    t_emb = OF._mlp_embed(p, "t_in", OF.timestep_embedding(t))
    p_emb = OF._mlp_embed(p, "p_in", pooled)
    torch.testing.assert_close(vec, 2 * t_emb + p_emb, rtol=1e-6, atol=1e-6)
    emb = OF.timestep_embedding(torch.tensor([0.03]))            # diffusers: guidance * 1000 -> sinusoid, cos first
    assert emb.shape == (1, 256) and float(emb[0, 0]) == pytest.approx(__import__("math").cos(30.0), abs=1e-5)
    assert float(emb[0, 128]) == pytest.approx(__import__("math").sin(30.0), abs=1e-5)
    cfg0 = OF.FluxConfig(in_channels=64, d=256, heads=2, n_double=1, n_single=1, txt_dim=48, pooled_dim=32, guidance=False)
    pz = dict(p)
    pz["g_in.w2"], pz["g_in.b2"] = torch.zeros_like(p["g_in.w2"]), torch.zeros_like(p["g_in.b2"])
    torch.testing.assert_close(OF.temb_vector(pz, cfg, t, torch.tensor([30.0, 2.5]), pooled),
                               OF.temb_vector(p, cfg0, t, None, pooled), rtol=1e-6, atol=1e-6)
