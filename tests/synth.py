"""Seeded synthetic inputs shared by the parity tests."""
import numpy as np


def corpus(n, d, seed, dtype=np.float32):
    """L2-normalised Gaussian rows, float32 [n,d]."""
    g = np.random.default_rng(seed)
    x = g.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(dtype)


def well_separated(n, d, nq, k, seed, min_gap=1e-6, rounds=50):
    """Corpus + queries whose top-(k+1) float64 scores are at least `min_gap` apart for every query
    (the fp32 kernel's summation error is ~1e-7), so that exact index parity is well-posed. Near
    ties are repaired by shrinking the lower-ranked row by 0.1 % (rows need not be exactly unit)."""
    x = corpus(n, d, seed)
    q = corpus(nq, d, seed + 1)
    kk = min(k + 1, n)
    if n <= 1:
        return x, q
    for _ in range(rounds):
        s = q.astype(np.float64) @ x.astype(np.float64).T
        bad = False
        for i in range(nq):
            top = np.argsort(-s[i], kind="stable")[:kk]
            gaps = -np.diff(s[i, top])
            for j in np.nonzero(gaps < min_gap)[0]:
                x[top[j + 1]] *= np.float32(0.999)
                bad = True
        if not bad:
            return x, q
    raise RuntimeError("could not generate a well separated corpus")
