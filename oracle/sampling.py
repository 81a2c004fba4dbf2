"""Sampling half of the oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates what the reference does at ``/root/reference/csm_mlx/generation.py:51-54,81-84``
(``argmax`` if temperature == 0 else ``mx.random.categorical(logits / T)``) and the ``mlx_lm.sample_utils``
filters its README/CLI pass in (``cli/generate.py:168-174``; un-vendored, pin mlx-lm>=0.22.0).

``mx.random.categorical`` is argmax(logits + Gumbel noise); MLX's own bit stream cannot be reproduced here, so
the noise source is DEFINED as Philox4x32-10 (Salmon et al., SC'11) with key = seed and counter
(index // 4, draw_lo, draw_hi, row), taking word index % 4; u = ((x >> 8) + 0.5) / 2^24; g = -log(-log(u)).
Filters (all on p = softmax(logits), one probability threshold, ties kept): top-k keeps the k largest;
top-p keeps the descending prefix whose exclusive cumulative mass is < p; min-p keeps p >= min_p * p_max but at
least the ``min_keep`` most likely.
"""

from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key) -> np.ndarray:
    """counter (..., 4) uint32, key (k0, k1) -> (..., 4) uint32."""
    c = counter.astype(np.uint64)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def gumbel(V: int, seed: int, draw: int, row: int) -> np.ndarray:
    idx = np.arange(V)
    ctr = np.zeros((V, 4), dtype=np.uint32)
    ctr[:, 0] = idx >> 2
    ctr[:, 1] = draw & 0xFFFFFFFF
    ctr[:, 2] = (draw >> 32) & 0xFFFFFFFF
    ctr[:, 3] = row
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    x = r[idx, idx & 3]
    u = ((x >> 8).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    return -np.log(-np.log(u, dtype=np.float32), dtype=np.float32)


def keep_mask(logits: np.ndarray, top_k: int = 0, top_p: float = 0.0, min_p: float = 0.0, min_keep: int = 1) -> np.ndarray:
    """Boolean mask of the tokens that survive the filters (logits (V,) fp32)."""
    V = logits.shape[0]
    e = np.exp((logits - logits.max()).astype(np.float32), dtype=np.float32)
    s = np.sort(e)[::-1]
    n1 = top_k if 0 < top_k < V else V
    t = s[n1 - 1]
    if 0.0 < top_p < 1.0:
        # nucleus on exact integer masses q = floor(e * 2^32) (order-independent sums: the definition every CUDA sampler of
        # the library shares): token i stays while the mass of the strictly more likely tokens is below top_p * total
        q = np.floor(s.astype(np.float64) * 4294967296.0).astype(np.uint64)
        need = float(np.float32(top_p)) * float(int(q.sum(dtype=np.uint64)))
        c = 0
        n2 = 0
        for i in range(n1):
            if float(c) < need:
                n2 = i + 1
            else:
                break
            c += int(q[i])
        t = max(t, s[n2 - 1])
    if min_p > 0.0:
        mk = min(max(min_keep, 1), V)
        t = max(t, min(np.float32(min_p) * s[0], s[mk - 1]))
    return e >= t


def sample(logits: np.ndarray, temperature: float, seed: int = 0, draw: int = 0, row: int = 0, top_k: int = 0,
           top_p: float = 0.0, min_p: float = 0.0, min_keep: int = 1) -> int:
    logits = np.asarray(logits, dtype=np.float32)
    if temperature == 0:
        return int(np.argmax(logits))  # first index on ties
    score = logits * np.float32(1.0 / temperature) + gumbel(logits.shape[0], seed, draw, row)
    if (0 < top_k < logits.shape[0]) or (0.0 < top_p < 1.0) or min_p > 0.0:
        score = np.where(keep_mask(logits, top_k, top_p, min_p, min_keep), score, -np.inf)
    return int(np.argmax(score))
