"""Oracle: exact inner-product search + style re-rank (numpy). TEST INFRASTRUCTURE ONLY.

Follows retrieval/clip100_resnet_style_all_shots.py:
  * clip_first_stage_retrieval :396-451 - np.vstack of the per-source feature arrays (:419), cast
    to float32 (:429), faiss.IndexFlatIP(d).add / .search(q, min(top_k, N)) (:425-434): all N inner
    products, top-k by score descending. faiss-cpu 1.10.0 is not installable here -> PARITY
    UNPINNED; faiss leaves the order of equal scores unspecified, this oracle (and the kernel)
    put the lower row id first.
  * resnet_second_stage_rerank :454-497 - np.linalg.norm(fq - fi) (:474), Python stable ascending
    sort on the distance (:485), similarity = 1/(1+d) (:492), rank = i+1 (:491). PINNED by
    tests/golden/rerank.json.
Scores are accumulated in float64 and rounded once to float32 (summation-order independent).
"""
from __future__ import annotations

import numpy as np

FAISS_MISSING_SCORE = np.float32(-3.4028234663852886e38)  # faiss fills missing slots with lowest()


def ip_scores(X: np.ndarray, q: np.ndarray) -> np.ndarray:
    """[N,d] x [nq,d] -> float32 [nq,N], accumulated in float64."""
    return (np.asarray(q, np.float64) @ np.asarray(X, np.float64).T).astype(np.float32)


def ip_topk(X: np.ndarray, q: np.ndarray, k: int, base_id: int = 0):
    """D float32 [nq,k], I int64 [nq,k]; score descending, ties -> lower id; (-FLT_MAX, -1) padding."""
    X = np.asarray(X, np.float32)
    q = np.asarray(q, np.float32)
    nq, n = q.shape[0], X.shape[0]
    D = np.full((nq, k), FAISS_MISSING_SCORE, np.float32)
    I = np.full((nq, k), -1, np.int64)
    if n == 0:
        return D, I
    S = ip_scores(X, q)
    ids = np.arange(n, dtype=np.int64)
    for i in range(nq):
        order = np.lexsort((ids, -S[i].astype(np.float64)))[:k]  # primary: score desc, then id asc
        D[i, : len(order)] = S[i, order]
        I[i, : len(order)] = order + base_id
    return D, I


def ip_topk_from_scores(S: np.ndarray, k: int, ids: np.ndarray | None = None):
    """Top-k of precomputed float32 scores [nq,N] with the same ordering rule."""
    nq, n = S.shape
    ids = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, np.int64)
    D = np.full((nq, k), FAISS_MISSING_SCORE, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for i in range(nq):
        order = np.lexsort((ids, -S[i].astype(np.float64)))[:k]
        D[i, : len(order)] = S[i, order]
        I[i, : len(order)] = ids[order]
    return D, I


def merge_topk(Dg: np.ndarray, Ig: np.ndarray, k: int):
    """Merge per-shard results [nq, lists, k_in] -> global top-k (score desc, id asc; id<0 = empty)."""
    nq = Dg.shape[0]
    D = np.full((nq, k), FAISS_MISSING_SCORE, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for i in range(nq):
        d = Dg[i].reshape(-1)
        ids = Ig[i].reshape(-1)
        valid = ids >= 0
        d, ids = d[valid], ids[valid]
        order = np.lexsort((ids, -d.astype(np.float64)))[:k]
        D[i, : len(order)] = d[order]
        I[i, : len(order)] = ids[order]
    return D, I


def sharded_ip_topk(X: np.ndarray, q: np.ndarray, k: int, bounds):
    """Reference semantics of the sharded search: per-shard top-k then merge."""
    Ds, Is = [], []
    for lo, hi in bounds:
        D, I = ip_topk(X[lo:hi], q, k, base_id=lo)
        Ds.append(D)
        Is.append(I)
    return merge_topk(np.stack(Ds, 1), np.stack(Is, 1), k)


def rerank_by_style(query_feat: np.ndarray, cand_feats, first_stage):
    """resnet_second_stage_rerank :454-497 given precomputed 128-d features.

    cand_feats[i] is None when the candidate image failed to load (silently dropped, :472).
    first_stage: list of dicts with similarity/image_path/source_dataset (the :437-445 records)."""
    rer = []
    for res, f in zip(first_stage, cand_feats):
        if f is None:
            continue
        dist = np.linalg.norm(np.asarray(query_feat, np.float32) - np.asarray(f, np.float32))
        rer.append({"clip_similarity": res["similarity"], "resnet_distance": float(dist),
                    "image_path": res["image_path"],
                    "source_dataset": res.get("source_dataset", "unknown")})
    rer.sort(key=lambda x: x["resnet_distance"])
    return [{"rank": i + 1, "similarity": float(1.0 / (1.0 + r["resnet_distance"])),
             "image_path": r["image_path"], "source_dataset": r["source_dataset"]}
            for i, r in enumerate(rer)]


def min_topk_gap(X: np.ndarray, q: np.ndarray, k: int) -> float:
    """Smallest gap between adjacent float64 scores among the top-(k+1) of each query: the synthetic
    generators assert this is well above fp32 rounding so that exact index parity is well-posed."""
    S = np.asarray(q, np.float64) @ np.asarray(X, np.float64).T
    gap = np.inf
    for i in range(S.shape[0]):
        top = np.sort(S[i])[::-1][: k + 1]
        if len(top) > 1:
            gap = min(gap, float(np.min(-np.diff(top))))
    return gap
