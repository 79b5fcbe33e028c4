"""Seeded synthetic inputs shared by the parity tests."""
import numpy as np


def corpus(n, d, seed, dtype=np.float32):
    """L2-normalised Gaussian rows, float32 [n,d]."""
    g = np.random.default_rng(seed)
    x = g.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(dtype)


def well_separated(n, d, nq, k, seed, min_gap=2e-6, tries=20):
    """Corpus + queries whose top-(k+1) float64 scores are at least `min_gap` apart, so that exact
    index parity between fp32 summation orders is well-posed."""
    from oracle.ip_topk import min_topk_gap
    for t in range(tries):
        x = corpus(n, d, seed + 1000 * t)
        q = corpus(nq, d, seed + 1000 * t + 1)
        if n <= 1 or min_topk_gap(x, q, min(k, n - 1)) >= min_gap:
            return x, q
    raise RuntimeError("could not generate a well separated corpus")
