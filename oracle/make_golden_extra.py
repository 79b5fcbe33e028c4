"""Second batch of golden vectors produced by running the REFERENCE's own functions (TEST INFRASTRUCTURE ONLY).

    python oracle/make_golden_extra.py        (authoring container only: needs /root/reference)

Writes tests/golden/ref_extra.json + ref_extra_arrays.npz:
  clean_paths     clean_image_path (retrieval/clip100_resnet_style_all_shots.py:77-86) on both rewrite prefixes
  dataset_params  get_strength_param / get_guidance_scale_param / get_image_prompt_scale_param /
                  get_upscale_dimension_param / get_redux_prompt (outpainting_updown_sampling_redux.py:1363-1381)
  first_stage     clip_first_stage_retrieval (:396-451) run as the reference wrote it - dict-order vstack, float32 cast,
                  record format, min(top_k, N) - over a seeded multi-source corpus. `faiss` itself is absent offline, so the
                  script's `faiss.IndexFlatIP` is served by a 12-line exact stand-in (float64 dot products rounded once to
                  float32, descending, ties -> lower id): the records pin the CALL SITE (ordering of sources, ids, record keys,
                  k clamp), not faiss's own last-ulp behaviour, which stays "parity unpinned".
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle.make_golden import GOLD, load_reference_outpaint, load_reference_retrieval  # noqa: E402


class _ExactFlatIP:
    """Stand-in for faiss.IndexFlatIP used ONLY to drive the reference's call site."""

    def __init__(self, d):
        self.d, self.x = d, np.zeros((0, d), np.float32)

    def add(self, x):
        assert x.dtype == np.float32 and x.shape[1] == self.d
        self.x = np.vstack([self.x, x])

    def search(self, q, k):
        s = (q.astype(np.float64) @ self.x.astype(np.float64).T).astype(np.float32)
        order = np.lexsort((np.arange(s.shape[1])[None].repeat(len(q), 0), -s), axis=1)[:, :k]
        return np.take_along_axis(s, order, 1), order.astype(np.int64)


def corpus(seed, n, d, dtype):
    g = np.random.default_rng(seed)
    x = g.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(dtype)


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        ref = load_reference_retrieval(tmp)
    cases = ["../../pipeline/datasets/DIOR/train/00011.jpg", "../../datasets/coco/train2017/000000000009.jpg",
             "../../pipeline/datasets/coco/train2017/1.jpg", "./coco/train2017/2.jpg", "/abs/datasets/coco/3.jpg",
             "../../datasets/cocoa/4.jpg", "../../datasets/ArTaxOr/5.jpg", "", None, 17]
    out["clean_paths"] = [{"in": c, "out": ref.clean_image_path(c)} for c in cases]

    ref.faiss.IndexFlatIP = _ExactFlatIP
    d = 64
    feats = {"coco": corpus(11, 300, d, np.float16), "empty": np.zeros((0, d), np.float32), "none": None,
             "mini_imagenet": corpus(12, 150, d, np.float32)}
    paths = {"coco": [f"../../datasets/coco/train2017/{i:012d}.jpg" for i in range(300)], "empty": [], "none": [],
             "mini_imagenet": [f"./mini/n{i:05d}.JPEG" for i in range(150)]}
    feats["mini_imagenet"][7] = feats["coco"][5].astype(np.float32)          # an exact duplicate across sources (tie)
    g = np.random.default_rng(13)
    queries = [feats["coco"][5].astype(np.float32) + 0.02 * g.standard_normal(d).astype(np.float32),
               feats["mini_imagenet"][40], g.standard_normal(d).astype(np.float32)]
    fs = []
    for qi, q in enumerate(queries):
        for top_k in (10, 100, 1000):
            recs = ref.clip_first_stage_retrieval(q, feats, paths, top_k=top_k)
            fs.append({"query": qi, "top_k": top_k, "records": recs})
    fs.append({"query": 0, "top_k": 5, "only_empty": True,
               "records": ref.clip_first_stage_retrieval(queries[0], {"empty": feats["empty"], "none": None},
                                                         {"empty": [], "none": []}, top_k=5)})
    out["first_stage"] = fs
    np.savez_compressed(GOLD / "ref_extra_arrays.npz", coco=feats["coco"], mini_imagenet=feats["mini_imagenet"],
                        queries=np.stack(queries))

    op = load_reference_outpaint()
    names = ["FISH", "DIOR", "ArTaxOr", "UODD", "NEU-DET", "clipart1k", "NWPU_VHR-10", "Camouflage", "coco", "unknown_ds"]
    out["dataset_params"] = {n: {"strength": op.get_strength_param(n), "guidance_scale": op.get_guidance_scale_param(n),
                                 "image_prompt_scale": op.get_image_prompt_scale_param(n),
                                 "upscale_dimension": op.get_upscale_dimension_param(n),
                                 "redux_prompt": op.get_redux_prompt(n)} for n in names}
    out["misc"] = {"create_gpu_process_id": op.create_gpu_process_id("20250101_abc", 3),
                   "extract_sample_prefix": [{"in": p, "out": op.extract_sample_prefix(p)} for p in
                                             ("a/b/00011_bbox0.jpg", "x.jpg", "/p/q/IMG_0001_2_3.png")]}
    json.dump(out, open(GOLD / "ref_extra.json", "w"), indent=1, ensure_ascii=False)
    print("written", GOLD / "ref_extra.json")


if __name__ == "__main__":
    main()
