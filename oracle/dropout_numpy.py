"""ORACLE — test infrastructure only (never imported by the product path).

Dropout masks of the B200 path, restated in numpy: the counter-based generator of simxns_b200/csrc/dropout.cuh
(Philox4x32 with 7 rounds; Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11 — constants
0xD2511F53 / 0xCD9E8D57, key increments 0x9E3779B9 / 0xBB67AE85).

What the reference does: nn.Dropout(p) with torch's own Philox stream at the HF sites
(/root/reference/LEAD/modeling_bert.py:239 embeddings, :358 attention probabilities, :386 / :464 hidden states after the
attention-output / FFN-output dense; DistilBERT /root/reference/LEAD/modeling_distilbert.py:133, :221, :254; p from
/root/reference/SimANS/model/models.py:66-72).  torch's stream depends on its kernel tiling and cannot be matched, so the
contract is: same sites, same Bernoulli(1-p) keep / (1/(1-p)) scaling semantics, masks given by THIS generator.  With the
masks fixed the forward/backward are deterministic and are checked exactly against oracle/bert_numpy.py.

  one call -> 4 x 32 bits -> 8 x 16-bit uniforms for columns [8*c0, 8*c0+8) of one row
  counter = (column >> 3, row, plane, site | layer << 8), key = (seed & 0xffffffff, seed >> 32)
  column e uses word e >> 1: low half for even e, high half for odd e; dropped iff uniform < round(p * 65536)
"""
from __future__ import annotations

import numpy as np

SITE_EMBED, SITE_ATTN, SITE_ATTN_OUT, SITE_FFN_OUT = 0, 1, 2, 3
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox7(c0, c1, c2, c3, k0, k1, rounds=7):
    """vectorised Philox4x32-R (R = 7 in the product; R = 10 reproduces the published known-answer vectors);
    counters are broadcastable integer arrays, keys python ints -> 4 uint32 arrays"""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3)])
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def threshold(p: float) -> int:
    return int(min(65535, max(0, int(p * 65536.0 + 0.5))))


def keep_scale(p: float) -> float:
    return 65536.0 / (65536.0 - threshold(p))


def uniforms16(seed: int, site: int, layer: int, rows, ncols: int, plane=0):
    """16-bit uniforms [len(rows), ncols] (ncols rounded up to 8 internally) of the given rows of one plane"""
    rows = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
    ng = (ncols + 7) // 8
    cg = np.arange(ng, dtype=np.uint64).reshape(1, -1)
    w = philox7(cg, rows, np.uint64(plane), np.uint64(site | (layer << 8)), seed & 0xFFFFFFFF, seed >> 32)
    out = np.empty((rows.shape[0], ng, 8), dtype=np.uint32)
    for i in range(4):
        out[:, :, 2 * i] = w[i] & np.uint32(0xFFFF)
        out[:, :, 2 * i + 1] = w[i] >> np.uint32(16)
    return out.reshape(rows.shape[0], ng * 8)[:, :ncols]


def hidden_mask(seed: int, site: int, layer: int, T: int, H: int, p: float):
    """[T, H] float64 multipliers (0 or 1/keep) of a hidden-state dropout site; row = token index b*S + s"""
    if p <= 0.0:
        return np.ones((T, H))
    u = uniforms16(seed, site, layer, np.arange(T), H)
    return np.where(u >= threshold(p), keep_scale(p), 0.0)


def attn_mask(seed: int, layer: int, B: int, heads: int, S: int, p: float):
    """[B, heads, S, S] multipliers of the attention-probability dropout; plane = b*heads + h, row = query, column = key"""
    if p <= 0.0:
        return np.ones((B, heads, S, S))
    out = np.empty((B, heads, S, S))
    thr, sc = threshold(p), keep_scale(p)
    for b in range(B):
        for h in range(heads):
            u = uniforms16(seed, SITE_ATTN, layer, np.arange(S), S, plane=b * heads + h)
            out[b, h] = np.where(u >= thr, sc, 0.0)
    return out
