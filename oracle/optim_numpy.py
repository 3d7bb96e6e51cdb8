"""ORACLE — test infrastructure only (never imported by the product path).

float64 numpy restatement of the optimizer tail of the reference train steps:

  clip_grad_norm : torch.nn.utils.clip_grad_norm_(params, max_norm)   call sites
                   /root/reference/SimANS/co_training/co_training_marco_train.py:246-259, LEAD/run_LEAD.py:232-238
                   total = ||(||g_1||, ..., ||g_n||)||_2 ; coef = min(1, max_norm / (total + 1e-6)) ; g *= coef
  adamw_torch    : torch.optim.AdamW, the optimizer of LEAD (/root/reference/LEAD/util.py:901)
  adamw_hf       : transformers.AdamW, the optimizer of the SimANS / PROD drivers
                   (/root/reference/SimANS/co_training/co_training_marco_train.py:22,57-69).  The class lives in the
                   un-vendored third-party package `transformers` (v4.x, optimization.py `AdamW.step`; no version is
                   pinned by the reference) and was REMOVED from the transformers 5.5 installed here, so
                   it cannot be executed in this image: its published algorithm is restated below —
                       exp_avg    = b1 exp_avg + (1-b1) g
                       exp_avg_sq = b2 exp_avg_sq + (1-b2) g^2
                       denom      = sqrt(exp_avg_sq) + eps
                       step_size  = lr * sqrt(1-b2^t) / (1-b1^t)      (lr when correct_bias=False)
                       p         -= step_size * exp_avg / denom
                       p         -= lr * weight_decay * p              (after the update, only if weight_decay > 0)

Pins: adamw_torch and clip_grad_norm are checked against the installed torch (tests/test_optim_cpu.py); adamw_hf is
"parity unpinned" by executable reference code (none exists here) — it is pinned only to the algebraic relation with
adamw_torch that the two published algorithms imply (eps = 0, weight_decay = 0 => identical updates).
"""
from __future__ import annotations

import numpy as np


def clip_grad_norm(grads, max_norm):
    """-> (total_norm, clipped grads).  `grads`: list of arrays (None entries are skipped and returned as None)."""
    sq = sum(float((np.asarray(g, np.float64) ** 2).sum()) for g in grads if g is not None)
    total = np.sqrt(sq)
    coef = min(1.0, max_norm / (total + 1e-6))
    return total, [None if g is None else np.asarray(g, np.float64) * coef for g in grads]


def adamw_torch(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
    """one torch.optim.AdamW update of a single tensor; `step` counts from 1.  -> (p, m, v)"""
    p, g, m, v = (np.asarray(x, np.float64) for x in (p, g, m, v))
    b1, b2 = betas
    p = p * (1.0 - lr * weight_decay)
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def adamw_hf(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
    """one transformers.AdamW (v4) update of a single tensor; `step` counts from 1.  -> (p, m, v)"""
    p, g, m, v = (np.asarray(x, np.float64) for x in (p, g, m, v))
    b1, b2 = betas
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    denom = np.sqrt(v) + eps
    step_size = lr
    if correct_bias:
        step_size = lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    p = p - step_size * m / denom
    if weight_decay > 0.0:
        p = p - lr * weight_decay * p
    return p, m, v
