"""Host logic of the retrieval entry point (no GPU): flag surface equals the reference's, query discovery,
corpus discovery, contact sheet. The end-to-end run on a synthetic tree is tests/test_retrieval_cli_gpu.py."""
import json

import numpy as np
from PIL import Image

from domain_rag_b200 import retrieval_cli as RC

# flags of the reference parser (retrieval/clip100_resnet_style_all_shots.py:967-996) with their defaults
REFERENCE_FLAGS = {
    "datasets": ["ArTaxOr", "DIOR", "FISH", "NEU-DET", "UODD", "clipart1k"], "shots": [1, 5, 10],
    "coco_dir": "./coco", "mini_imagenet_dir": "./miniimagenet", "dataset_source": "coco", "clip_top_k": 100,
    "output_dir": None, "gpu_id": 0, "pretrained_coco_features": "./coco_embeddings_global.pt",
    "pretrained_coco_paths": None, "pretrained_mini_imagenet_features": None,
    "pretrained_mini_imagenet_paths": None, "global_features": False, "force_recompute": False,
    "lamainpaint_dir": None, "force_recompute_inpainted": False,
}


def make_tree(root, n_coco=6, samples=("a1", "b2", "c3")):
    g = np.random.default_rng(0)
    (root / "coco" / "train2017" / "sub").mkdir(parents=True)
    for i in range(n_coco):
        arr = (g.random((40 + i, 50, 3)) * 255).astype(np.uint8)
        ext = ("jpg", "png", "jpeg")[i % 3]
        sub = "sub/" if i % 2 else ""
        Image.fromarray(arr).save(root / "coco" / "train2017" / f"{sub}im{i}.{ext}")
    shot = root / "lamainpaint" / "DS" / "1_shot"
    shot.mkdir(parents=True)
    for s in samples:
        Image.fromarray((g.random((32, 48, 3)) * 255).astype(np.uint8)).save(shot / f"{s}.jpg")
    return shot


def test_flag_surface_matches_reference():
    ns = vars(RC.build_parser().parse_args([]))
    for k, v in REFERENCE_FLAGS.items():
        assert ns[k] == v, k
    ns = RC.build_parser().parse_args(["--datasets", "DIOR", "--shots", "5", "--dataset-source", "both",
                                       "--clip-top-k", "50", "--force-recompute"])
    assert ns.datasets == ["DIOR"] and ns.shots == [5] and ns.dataset_source == "both" and ns.clip_top_k == 50


def test_query_and_corpus_discovery(tmp_path):
    shot = make_tree(tmp_path)
    s2i, s2c = RC.get_inpainted_images("DS", 1, str(tmp_path / "lamainpaint"))
    assert sorted(s2i) == ["a1", "b2", "c3"] and s2c == {"a1": "a1", "b2": "b2", "c3": "c3"}
    json.dump({"a1": "beetle", "zz": "unused"}, open(shot / "category_mapping.json", "w"))
    _, s2c = RC.get_inpainted_images("DS", 1, str(tmp_path / "lamainpaint"))
    assert s2c == {"a1": "beetle", "b2": "b2", "c3": "c3"}
    assert RC.get_inpainted_images("DS", 5, str(tmp_path / "lamainpaint")) == ({}, {})
    imgs = RC.list_corpus_images("coco", str(tmp_path / "coco"))
    assert len(imgs) == 6 and any("sub" in p for p in imgs)
    assert RC.list_corpus_images("coco", str(tmp_path / "nope")) == []
    assert RC.list_corpus_images("mini-imagenet", str(tmp_path / "coco")) == []


def test_contact_sheet(tmp_path):
    make_tree(tmp_path)
    imgs = RC.list_corpus_images("coco", str(tmp_path / "coco"))
    out = tmp_path / "v.jpg"
    RC.visualize_results(imgs[0], imgs[1:] + [str(tmp_path / "missing.png")], str(out))
    assert Image.open(out).size == (1024, 3 * 256 + 54)
