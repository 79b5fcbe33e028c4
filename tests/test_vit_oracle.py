"""CPU: the CLIP ViT oracle against transformers.CLIPVisionModelWithProjection (independent code for the
same architecture; OpenAI's own package is not installable offline)."""
import pytest
import torch

from oracle import vit as OV


def to_hf(state, cfg):
    sd = {"vision_model.embeddings.class_embedding": state["visual.class_embedding"],
          "vision_model.embeddings.patch_embedding.weight": state["visual.conv1.weight"],
          "vision_model.embeddings.position_embedding.weight": state["visual.positional_embedding"],
          "vision_model.pre_layrnorm.weight": state["visual.ln_pre.weight"],
          "vision_model.pre_layrnorm.bias": state["visual.ln_pre.bias"],
          "vision_model.post_layernorm.weight": state["visual.ln_post.weight"],
          "vision_model.post_layernorm.bias": state["visual.ln_post.bias"],
          "visual_projection.weight": state["visual.proj"].t().contiguous()}
    w = cfg.width
    for i in range(cfg.layers):
        p, h = f"visual.transformer.resblocks.{i}.", f"vision_model.encoder.layers.{i}."
        W, b = state[p + "attn.in_proj_weight"], state[p + "attn.in_proj_bias"]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            sd[h + f"self_attn.{n}.weight"], sd[h + f"self_attn.{n}.bias"] = W[j * w:(j + 1) * w], b[j * w:(j + 1) * w]
        sd[h + "self_attn.out_proj.weight"], sd[h + "self_attn.out_proj.bias"] = state[p + "attn.out_proj.weight"], state[p + "attn.out_proj.bias"]
        sd[h + "layer_norm1.weight"], sd[h + "layer_norm1.bias"] = state[p + "ln_1.weight"], state[p + "ln_1.bias"]
        sd[h + "layer_norm2.weight"], sd[h + "layer_norm2.bias"] = state[p + "ln_2.weight"], state[p + "ln_2.bias"]
        sd[h + "mlp.fc1.weight"], sd[h + "mlp.fc1.bias"] = state[p + "mlp.c_fc.weight"], state[p + "mlp.c_fc.bias"]
        sd[h + "mlp.fc2.weight"], sd[h + "mlp.fc2.bias"] = state[p + "mlp.c_proj.weight"], state[p + "mlp.c_proj.bias"]
    return sd


@pytest.mark.parametrize("cfg", [OV.ViTConfig(128, 2, 2, 32, 224, 64), OV.ViTConfig(192, 3, 3, 14, 224, 96)])
def test_oracle_matches_transformers_clip(cfg):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    state = OV.init_state(cfg, seed=7)
    hf_cfg = CLIPVisionConfig(hidden_size=cfg.width, intermediate_size=4 * cfg.width, projection_dim=cfg.out_dim,
                              num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads, image_size=cfg.image,
                              patch_size=cfg.patch, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    m = CLIPVisionModelWithProjection(hf_cfg).eval()
    missing, unexpected = m.load_state_dict(to_hf(state, cfg), strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = m(pixel_values=x).image_embeds
        got = OV.encode_image(state, cfg, x)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def test_default_configs_match_reference_model_names():
    assert OV.CONFIGS["ViT-B/32"].tokens == 50 and OV.CONFIGS["ViT-B/32"].out_dim == 512
    assert OV.CONFIGS["ViT-L/14"].tokens == 257 and OV.CONFIGS["ViT-L/14"].out_dim == 768
    from domain_rag_b200 import clip as C
    for n in OV.CONFIGS:
        a, b = OV.CONFIGS[n], C.CONFIGS[n]
        assert (a.width, a.layers, a.heads, a.patch, a.image, a.out_dim) == (b.width, b.layers, b.heads, b.patch, b.image, b.out_dim)
    s1, s2 = OV.init_state(OV.CONFIGS["ViT-B/32"], 5), C.random_state(C.CONFIGS["ViT-B/32"], 5)
    assert all(torch.equal(s1[k], s2[k]) for k in s1)
