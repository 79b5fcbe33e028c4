"""Host-side mirror of the reference's two-stage retrieval functions
(retrieval/clip100_resnet_style_all_shots.py), same names / argument meaning / return records,
with the third-party calls served by libdomainrag_b200.so:

  clip_first_stage_retrieval  :396-451  faiss.IndexFlatIP add/search -> IndexFlatIP (HBM resident,
                                        cached across queries instead of rebuilt per query)
  compute_resnet_features     :180-203  cv2 read/resize -> fused stem+stats kernel
  resnet_second_stage_rerank  :454-497  101 batch-1 launches -> ONE batched launch; distance /
                                        stable sort / 1/(1+d) on the host exactly as the reference
Error convention is the reference's: print + return []/None/first-stage results and continue.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .index import IndexFlatIP

_INDEX_CACHE: dict = {}
# path -> 128-d style vector of corpus images already seen: the reference recomputes the statistics of all
# 100 candidates for every query (:468-470) although popular corpus images recur across queries.
# The cache lives on the encoder object (a style vector is only valid for the stem weights that produced it).
_STYLE_CACHE_MAX = 200_000


def _style_cache(resnet_model) -> Dict[str, np.ndarray]:
    cache = getattr(resnet_model, "_style_cache", None)
    if cache is None:
        cache = {}
        try:
            resnet_model._style_cache = cache
        except AttributeError:      # foreign encoder objects without a __dict__: no caching
            pass
    if len(cache) > _STYLE_CACHE_MAX:
        cache.clear()
    return cache


def clean_image_path(path):
    """reference :77-86 - two rewrites, the second only when the first did not apply: the stray "pipeline/" prefix some
    caches carry is stripped; otherwise corpus paths cached as ../../datasets/coco/... are redirected to ./coco/...
    (the default --pretrained-coco-features path lists rely on it)."""
    if not isinstance(path, str):
        return path
    for stale, current in (("../../pipeline/datasets", "../../datasets"), ("../../datasets/coco", "./coco")):
        if stale in path:
            return path.replace(stale, current)
    return path


def _cached_index(dataset_features: Dict[str, np.ndarray], d: int, device: int) -> IndexFlatIP:
    """One resident index per (set of feature arrays); the reference re-adds N*d floats per query."""
    # content fingerprint (17 sampled rows), not id(): a recomputed feature array may reuse a freed array's address
    def fingerprint(f):
        a = np.asarray(f)
        rows = a[:: max(1, len(a) // 16)][:17]
        return hash((a.shape, str(a.dtype), np.ascontiguousarray(rows).tobytes()))

    key = (device, d) + tuple((name, fingerprint(f)) for name, f in dataset_features.items()
                              if f is not None and len(f) > 0)
    ix = _INDEX_CACHE.get(key)
    if ix is None:
        _INDEX_CACHE.clear()  # keep a single corpus resident
        ix = IndexFlatIP(d, device)
        for name, f in dataset_features.items():
            if f is not None and len(f) > 0:
                ix.add(np.asarray(f, dtype=np.float32))  # dict order == reference vstack order (:406-419)
        _INDEX_CACHE[key] = ix
    return ix


def clip_first_stage_retrieval(query_feature, dataset_features, dataset_paths, top_k=100, device: int = 0):
    """Stage A: exact inner-product top-k over every source dataset; returns the reference records
    {similarity, image_path, source_dataset, index} in descending score."""
    all_paths: List[str] = []
    all_sources: List[str] = []
    n_total = 0
    for name, feats in dataset_features.items():
        if feats is not None and len(feats) > 0:
            paths = dataset_paths[name]
            print(f"将从{name}数据集({len(feats)}张图像)中检索")
            all_paths.extend(paths)
            all_sources.extend([name] * len(paths))
            n_total += len(feats)
    if n_total == 0:
        print("错误：没有可用的数据集特征")
        return []
    print(f"总共使用 {n_total} 张图像进行检索")
    try:
        q = np.asarray([query_feature], dtype=np.float32)
        ix = _cached_index(dataset_features, q.shape[1], device)
        D, I = ix.search(q, min(top_k, n_total))
        results = []
        for i, idx in enumerate(I[0]):
            if 0 <= idx < len(all_paths):
                results.append({"similarity": float(D[0][i]), "image_path": all_paths[idx],
                                "source_dataset": all_sources[idx], "index": int(idx)})
        return results
    except Exception as e:  # reference :449-451
        print(f"CLIP检索时出错: {e}")
        return []


def load_style_input(image_path: str):
    """reference :186-193: cv2.imread -> RGB -> resize(256,256) -> CHW. Returned as the uint8 pixels (host tensor): the
    float() / 255.0 of :193 runs inside the stem kernel's loader (same IEEE fp32 division), a quarter of the H2D bytes."""
    import cv2
    import torch
    img = cv2.imread(clean_image_path(image_path))
    if img is None:
        print(f"警告：无法读取图像 {image_path}")
        return None
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    img = cv2.resize(img, (256, 256))
    return torch.from_numpy(img).permute(2, 0, 1).contiguous()


def compute_resnet_features(image_path, model, device):
    """reference :180-203 - one image -> float32 [128] = cat(mean, std), or None on failure."""
    try:
        x = load_style_input(image_path)
        if x is None:
            return None
        return model.style_features(x.unsqueeze(0).to(device))[0].cpu().numpy()
    except Exception as e:
        print(f"计算ResNet特征时出错: {e}, 图像: {image_path}")
        return None


def compute_resnet_features_batch(image_paths: Sequence[str], model, device) -> List[Optional[np.ndarray]]:
    """All candidates of a query in ONE kernel launch (the reference does 101 batch-1 launches)."""
    import torch
    xs, slots = [], []
    for i, p in enumerate(image_paths):
        x = load_style_input(p)
        if x is not None:
            xs.append(x)
            slots.append(i)
    out: List[Optional[np.ndarray]] = [None] * len(image_paths)
    if xs:
        batch = torch.stack(xs).pin_memory().to(device, non_blocking=True)
        feats = model.style_features(batch).cpu().numpy()
        for j, i in enumerate(slots):
            out[i] = feats[j]
    return out


def rerank_by_style(query_feat, cand_feats, first_stage_results):
    """Distance / stable sort / similarity of reference :474-495 on precomputed 128-d features."""
    rerank_results = []
    for result, f in zip(first_stage_results, cand_feats):
        if f is None:
            continue  # reference :472 silently drops unreadable candidates
        distance = np.linalg.norm(np.asarray(query_feat, np.float32) - np.asarray(f, np.float32))
        rerank_results.append({"clip_similarity": result["similarity"], "resnet_distance": float(distance),
                               "image_path": clean_image_path(result["image_path"]),
                               "source_dataset": result.get("source_dataset", "unknown")})
    rerank_results.sort(key=lambda x: x["resnet_distance"])
    return [{"rank": i + 1, "similarity": float(1.0 / (1.0 + r["resnet_distance"])),
             "image_path": r["image_path"], "source_dataset": r["source_dataset"]}
            for i, r in enumerate(rerank_results)]


def resnet_second_stage_rerank(query_image_path, first_stage_results, resnet_model, device):
    """Stage B: re-rank the CLIP candidates by L2 distance of stem style statistics. One batched launch for the query and
    the candidates not yet in the style cache."""
    query_image_path = clean_image_path(query_image_path)
    cand_paths = [clean_image_path(r["image_path"]) for r in first_stage_results]
    cache = _style_cache(resnet_model)
    todo = [query_image_path] + [p for p in dict.fromkeys(cand_paths) if p not in cache]
    feats = compute_resnet_features_batch(todo, resnet_model, device)
    if feats[0] is None:
        print(f"警告：无法计算查询图像的ResNet特征: {query_image_path}")
        return first_stage_results
    for p, f in zip(todo[1:], feats[1:]):
        if f is not None:
            cache[p] = f
    return rerank_by_style(feats[0], [cache.get(p) for p in cand_paths], first_stage_results)
