"""GPU: the drop-in retrieval entry point end to end on a synthetic dataset tree - file surface of the
reference and agreement of every stage with the CPU oracle on the same inputs."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ip_topk as OT
from oracle import stem as OS
from tests.test_retrieval_cli import make_tree

pytestmark = pytest.mark.gpu


def test_end_to_end_tree(lib, tmp_path):
    from domain_rag_b200 import retrieval_cli as RC
    from domain_rag_b200.resnet import random_stem_state
    shot = make_tree(tmp_path, n_coco=40)
    json.dump({"a1": "beetle"}, open(shot / "category_mapping.json", "w"))
    out = tmp_path / "res"
    base = ["--coco-dir", str(tmp_path / "coco"), "--lamainpaint-dir", str(tmp_path / "lamainpaint"), "--output-dir", str(out),
            "--pretrained-coco-features", str(tmp_path / "none.pt"), "--clip-top-k", "10"]
    # no checkpoints and no explicit opt-in: refuse (the reference downloads pretrained CLIP / ResNet-50 here)
    assert RC.main(["--datasets", "DS", "--shots", "1", *base]) == 2
    assert not (out / "coco_clip_features.npy").exists()
    rc = RC.main(["--datasets", "DS", "MISSING", "--shots", "1", "5", *base, "--allow-random-init"])
    assert rc == 0
    assert json.load(open(out / "coco_clip_features.npy.meta.json")) == {"clip_weights": "ViT-B/32:random-init"}
    # file surface
    names = set(os.listdir(out))
    assert {"coco_clip_features.npy", "coco_image_paths.json", "DS_1_shot_inpainted_clip_features.npy",
            "DS_1_shot_inpainted_image_paths.json", "DS_1_shot_retrieval_results.json",
            "all_shots_retrieval_results.json", "DS_1_shot_beetle_a1_retrieval_results.json",
            "DS_1_shot_beetle_a1_visual.jpg", "DS_1_shot_b2_b2_retrieval_results.json"} <= names
    allr = json.load(open(out / "all_shots_retrieval_results.json"))
    assert set(allr) == {"DS", "MISSING"} and allr["MISSING"] == {} and set(allr["DS"]) == {"1_shot"}
    per = allr["DS"]["1_shot"]
    assert set(per) == {"beetle", "b2", "c3"}
    rec = per["beetle"][0]
    assert set(rec) == {"sample_id", "image_path", "category", "similar_images"} and rec["sample_id"] == "a1"
    sim = rec["similar_images"]
    assert len(sim) == 10 and [r["rank"] for r in sim] == list(range(1, 11))
    assert set(sim[0]) == {"rank", "similarity", "image_path", "source_dataset"}
    assert all(sim[i]["similarity"] >= sim[i + 1]["similarity"] for i in range(9))

    # stage parity on the same inputs: scan on the cached GPU embeddings vs the oracle scan, then style re-rank
    X = np.load(out / "coco_clip_features.npy")
    paths = json.load(open(out / "coco_image_paths.json"))
    Q = np.load(out / "DS_1_shot_inpainted_clip_features.npy")
    qpaths = json.load(open(out / "DS_1_shot_inpainted_image_paths.json"))
    assert X.dtype == np.float32 and X.shape == (40, 512) and np.allclose(np.linalg.norm(X, axis=1), 1, atol=1e-3)
    qi = [i for i, p in enumerate(qpaths) if p.endswith("a1.jpg")][0]
    D, I = OT.ip_topk(X, Q[qi:qi + 1], 10)
    first = [{"similarity": float(D[0][j]), "image_path": paths[i], "source_dataset": "coco", "index": int(i)}
             for j, i in enumerate(I[0])]
    import cv2
    state = random_stem_state(2000)

    def style(p):
        img = cv2.resize(cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB), (256, 256))
        return OS.style_features(torch.tensor(img).float().permute(2, 0, 1).unsqueeze(0) / 255.0, state).numpy()[0]

    want = OT.rerank_by_style(style(rec["image_path"]), [style(r["image_path"]) for r in first], first)
    assert [r["image_path"] for r in sim] == [r["image_path"] for r in want]
    np.testing.assert_allclose([r["similarity"] for r in sim], [r["similarity"] for r in want], rtol=1e-4)

    # second run is served from the caches and reproduces the same results bit for bit
    before = open(out / "DS_1_shot_retrieval_results.json").read()
    assert RC.main(["--datasets", "DS", "--shots", "1", *base, "--no-visual", "--allow-random-init"]) == 0
    assert open(out / "DS_1_shot_retrieval_results.json").read() == before

    # a run with weight FILES (here: differently seeded stand-ins in the OpenAI / torchvision key layouts) must not reuse
    # the random-init caches: tags differ -> features recomputed, stem weights taken from the file
    from domain_rag_b200 import clip
    torch.save(clip.random_state(clip.CONFIGS["ViT-B/32"], seed=7), tmp_path / "clip.pt")
    torch.save({**random_stem_state(9), "fc.weight": torch.zeros(3, 3)}, tmp_path / "resnet50.pt")
    assert RC.main(["--datasets", "DS", "--shots", "1", *base, "--no-visual", "--clip-weights", str(tmp_path / "clip.pt"),
                    "--resnet-weights", str(tmp_path / "resnet50.pt")]) == 0
    tag = json.load(open(out / "coco_clip_features.npy.meta.json"))["clip_weights"]
    assert tag.startswith("ViT-B/32:") and not tag.endswith("random-init")
    X2 = np.load(out / "coco_clip_features.npy")
    assert X2.shape == X.shape and np.abs(X2 - X).max() > 1e-2
    assert open(out / "DS_1_shot_retrieval_results.json").read() != before
    torch.save({"conv1.weight": torch.zeros(64, 3, 7, 7)}, tmp_path / "bad.pt")
    assert RC.main(["--datasets", "DS", "--shots", "1", *base, "--clip-weights", str(tmp_path / "clip.pt"),
                    "--resnet-weights", str(tmp_path / "bad.pt")]) == 2
