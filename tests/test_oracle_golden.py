"""CPU: the oracle against the golden vectors produced by the reference's own functions
(oracle/make_golden.py) and against its own invariants."""
import json

import numpy as np
import pytest
import torch

from domain_rag_b200.resnet import fold_stem, random_stem_state
from oracle import host_helpers, ip_topk, stem


def test_stem_oracle_matches_reference_golden(golden_dir):
    g = np.load(golden_dir / "stem_stats.npz")
    x = torch.rand(4, 3, 256, 256, generator=torch.Generator().manual_seed(int(g["input_seed"])))
    got = stem.style_features(x, random_stem_state(2000)).numpy()
    np.testing.assert_allclose(got, g["stats"], rtol=1e-5, atol=1e-6)


def test_stem_oracle_file_path_matches_reference_golden(golden_dir):
    import cv2
    g = np.load(golden_dir / "stem_stats.npz")
    state = random_stem_state(2000)
    for name, want in zip(g["file_names"], g["file_feats"]):
        img = cv2.cvtColor(cv2.imread(str(golden_dir / "images" / str(name))), cv2.COLOR_BGR2RGB)
        img = cv2.resize(img, (256, 256))
        t = torch.tensor(img).float().permute(2, 0, 1).unsqueeze(0) / 255.0
        got = stem.style_features(t, state).numpy()[0]
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_folded_stem_equals_unfolded():
    state = random_stem_state(2000)
    w, b = fold_stem(state["conv1.weight"], state["bn1.weight"], state["bn1.bias"],
                     state["bn1.running_mean"], state["bn1.running_var"])
    x = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    ref = stem.stem_forward(x, state)
    y = torch.nn.functional.conv2d(x, w, b, stride=2, padding=3).relu()
    y = torch.nn.functional.max_pool2d(y, 3, 2, 1)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-5)


def test_rerank_oracle_matches_reference_golden(golden_dir):
    g = json.load(open(golden_dir / "rerank.json"))
    s = np.load(golden_dir / "stem_stats.npz")
    feats = {str(n): f for n, f in zip(s["file_names"], s["file_feats"])}
    cand = [feats.get(r["image_path"]) for r in g["first_stage"]]
    got = ip_topk.rerank_by_style(feats[g["query"]], cand, g["first_stage"])
    assert [r["image_path"] for r in got] == [r["image_path"] for r in g["reranked"]]
    assert [r["rank"] for r in got] == [r["rank"] for r in g["reranked"]]
    np.testing.assert_allclose([r["similarity"] for r in got], [r["similarity"] for r in g["reranked"]],
                               rtol=1e-6)


def test_host_helpers_match_reference_golden(golden_dir):
    from PIL import Image
    g = json.load(open(golden_dir / "host_helpers.json"))
    arrs = np.load(golden_dir / "host_helpers_arrays.npz")
    for c in g["split"]:
        assert host_helpers.split_samples_for_gpus([f"s{i}" for i in range(c["n"])], c["gpus"]) == c["out"]
    for c in g["resolution"]:
        im = Image.new("RGB", tuple(c["size"]))
        if c.get("error"):
            with pytest.raises(ValueError):
                host_helpers.process_image_resolution(im)
            continue
        out, up, down, nu, nd = host_helpers.process_image_resolution(im)
        assert list(out.size) == c["out_size"] and (nu, nd) == (c["need_up"], c["need_down"])
        assert up == pytest.approx(c["up"]) and down == pytest.approx(c["down"])
    for c in g["mask"]:
        m = host_helpers.generate_outpaint_mask(tuple(c["size"]), [tuple(b) for b in c["boxes"]])
        np.testing.assert_array_equal(np.array(m), arrs[c["key"]])


def test_ip_topk_oracle_properties():
    g = np.random.default_rng(0)
    x = g.standard_normal((500, 32)).astype(np.float32)
    q = g.standard_normal((3, 32)).astype(np.float32)
    D, I = ip_topk.ip_topk(x, q, 10)
    s = ip_topk.ip_scores(x, q)
    for i in range(3):
        assert np.all(np.diff(D[i]) <= 0)
        assert set(I[i]) == set(np.argsort(-s[i], kind="stable")[:10])
        np.testing.assert_array_equal(D[i], s[i, I[i]])
    # ties: duplicated rows -> lower id first
    x2 = np.concatenate([x[:5], x[:5]], 0)
    D2, I2 = ip_topk.ip_topk(x2, x[:1], 4)
    assert list(I2[0][:2]) == [0, 5]
    # k > N pads with (-FLT_MAX, -1)
    D3, I3 = ip_topk.ip_topk(x[:3], q, 5)
    assert np.all(I3[:, 3:] == -1) and np.all(D3[:, 3:] == ip_topk.FAISS_MISSING_SCORE)
    # sharded == single
    from domain_rag_b200.index import shard_bounds
    Ds, Is = ip_topk.sharded_ip_topk(x, q, 10, shard_bounds(500, 3))
    np.testing.assert_array_equal(Is, I)
    np.testing.assert_array_equal(Ds, D)


# ------------------------------------------------------------------ second batch of reference-generated vectors
@pytest.fixture(scope="module")
def ref_extra(golden_dir):
    return json.load(open(golden_dir / "ref_extra.json")), np.load(golden_dir / "ref_extra_arrays.npz")


def test_clean_image_path_matches_reference_golden(ref_extra):
    """Both rewrites of retrieval/clip100_resnet_style_all_shots.py:77-86 (pipeline/ prefix, ../../datasets/coco -> ./coco)."""
    from domain_rag_b200.retrieval import clean_image_path
    for c in ref_extra[0]["clean_paths"]:
        assert clean_image_path(c["in"]) == c["out"], c


def test_dataset_tables_match_reference_golden(ref_extra):
    from domain_rag_b200 import hostlogic as H
    for name, want in ref_extra[0]["dataset_params"].items():
        p = H.dataset_params(name)
        assert dict(strength=p.strength, guidance_scale=p.guidance_scale, image_prompt_scale=p.image_prompt_scale,
                    upscale_dimension=p.upscale_dimension, redux_prompt=p.redux_prompt) == want, name


def test_first_stage_oracle_matches_reference_call_site(ref_extra):
    """oracle.ip_topk over the vstack of the non-empty sources (dict order, float32 cast) reproduces the records the
    reference's clip_first_stage_retrieval emitted (ids, order, scores, k clamp, duplicate across sources)."""
    g, arrs = ref_extra
    x = np.vstack([arrs["coco"].astype(np.float32), arrs["mini_imagenet"].astype(np.float32)])
    for case in g["first_stage"]:
        if case.get("only_empty"):
            assert case["records"] == []
            continue
        q = arrs["queries"][case["query"]][None].astype(np.float32)
        D, I = ip_topk.ip_topk(x, q, min(case["top_k"], len(x)))
        assert [r["index"] for r in case["records"]] == I[0].tolist()
        np.testing.assert_allclose([r["similarity"] for r in case["records"]], D[0], rtol=0, atol=1e-6)
