"""ORACLE — test infrastructure only (never imported by the product path).

float64 numpy restatement, with closed-form gradients, of the score/loss tail of the reference train steps:

  nll_global : BiEncoderNllLoss.calc            /root/reference/SimANS/model/models.py:468-505
               (called on all-gathered embeddings: MASTER/finetune/MS/run_de_model.py:180-234,
                LEAD/util.py:416-451 get_loss_dual — positives at pos_idx[q])
  kl_local   : SimANS retriever step            /root/reference/SimANS/co_training/co_training_marco_train.py:199-214
               KLDivLoss(batchmean)(log(softmax(s)+eps), softmax(teacher/T))
  ce_local   : -sum p_t log(p_s+eps) / B        /root/reference/SimANS/wiki/co_training_wiki_train.py:221-227 (adv_lambda=0),
                                                /root/reference/MASTER/finetune/MS/co_training_model.py:249-269,
                                                PROD `prob_loss` /root/reference/PROD/ProD_KD/model/models.py:735-740
  kd_local   : CrossBERTKDLoss.kd_loss KD_softmax /root/reference/PROD/ProD_KD/model/models.py:772-777
  nll_local  : CrossBERTKDLoss hard loss (positive = column 0)  .../models.py:694-712

Pinned against the reference's own code by oracle/make_golden.py -> tests/golden/loss_*.npz.
"""
from __future__ import annotations

import numpy as np


def _log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


def score_loss(Q, C, *, terms, q_begin=0, q_count=None, c_begin=0, group=1, pos_idx=None, teacher=None,
               w=None, temperature=1.0, eps=1e-7, score_scale=1.0, grad_scale=1.0):
    """Returns (losses dict, dQ_local [q_count,H], dC_local [q_count*group,H]).

    terms: iterable of {'nll_global','kl_local','ce_local','kd_local','nll_local'}; w: dict term -> weight (default 1).
    """
    Q = np.asarray(Q, np.float64)
    C = np.asarray(C, np.float64)
    GQ, H = Q.shape
    if q_count is None:
        q_count = GQ
    c_count = q_count * group
    w = dict(w or {})
    dQ = np.zeros((q_count, H))
    dC = np.zeros((c_count, H))
    out = dict(total=0.0, nll_global=0.0, kl_local=0.0, ce_local=0.0, kd_local=0.0, nll_local=0.0, n_correct=0.0)
    terms = set(terms)

    if "nll_global" in terms:
        S = Q @ C.T
        lsm = _log_softmax(S)
        pos = np.asarray(pos_idx, np.int64)
        loss = -lsm[np.arange(GQ), pos].mean()
        out["nll_global"] = loss
        out["n_correct"] = float((lsm.argmax(1) == pos).sum())
        wt = w.get("nll_global", 1.0)
        out["total"] += wt * loss
        dS = np.exp(lsm)
        dS[np.arange(GQ), pos] -= 1.0
        dS *= wt / GQ
        # gradients reach only the local rows / columns (remote slices are constants: LEAD/util.py:200-204)
        dQ += dS[q_begin:q_begin + q_count] @ C
        dC += dS[:, c_begin:c_begin + c_count].T @ Q

    local_terms = terms & {"kl_local", "ce_local", "kd_local", "nll_local"}
    if local_terms:
        Ql = Q[q_begin:q_begin + q_count]
        Cl = C[c_begin:c_begin + c_count].reshape(q_count, group, H)
        s = np.einsum("bh,bdh->bd", Ql, Cl) * score_scale
        B = q_count
        ds = np.zeros_like(s)
        lsm = _log_softmax(s)
        p = np.exp(lsm)
        if teacher is not None:
            tl = np.asarray(teacher, np.float64) / temperature
            t = np.exp(_log_softmax(tl))
        if "kl_local" in terms or "ce_local" in terms:
            logp = np.log(p + eps)
            r = t * p / (p + eps)
            g = -(r - p * r.sum(-1, keepdims=True)) / B
            if "kl_local" in terms:
                with np.errstate(divide="ignore", invalid="ignore"):
                    tlogt = np.where(t > 0, t * np.log(t), 0.0)
                loss = (tlogt - t * logp).sum() / B
                out["kl_local"] = loss
                out["total"] += w.get("kl_local", 1.0) * loss
                ds += w.get("kl_local", 1.0) * g
            if "ce_local" in terms:
                loss = -(t * logp).sum() / B
                out["ce_local"] = loss
                out["total"] += w.get("ce_local", 1.0) * loss
                ds += w.get("ce_local", 1.0) * g
        if "kd_local" in terms:
            T = temperature
            lsT = _log_softmax(s / T)
            with np.errstate(divide="ignore", invalid="ignore"):
                kl = np.where(t > 0, t * (np.log(t) - lsT), 0.0)
            loss = kl.sum(1).mean() * T * T
            out["kd_local"] = loss
            out["total"] += w.get("kd_local", 1.0) * loss
            ds += w.get("kd_local", 1.0) * (T / B) * (np.exp(lsT) - t)
        if "nll_local" in terms:
            loss = -lsm[:, 0].mean()
            out["nll_local"] = loss
            out["total"] += w.get("nll_local", 1.0) * loss
            g = p.copy()
            g[:, 0] -= 1.0
            ds += w.get("nll_local", 1.0) * g / B
        ds *= score_scale
        dQ += np.einsum("bd,bdh->bh", ds, Cl)
        dC += (ds[:, :, None] * Ql[:, None, :]).reshape(c_count, H)

    return out, dQ * grad_scale, dC * grad_scale
