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
  mse_local / mse_b2a : kd_loss KD_logit, 0.5 MSELoss(mean)(s, t)   .../models.py:778-780, 1057-1059
  dkd_local / dkd_b2a : dkd_loss (decoupled KD: alpha TCKD + beta NCKD)  .../models.py:783-812, 1070-1094
  embed_mse  : BiEncoderKDLoss.bi_logit, 0.5 MSE(q_s, q_t) + 0.5 MSE(c_s, c_t)   .../models.py:1062-1067

Pinned against the reference's own code by oracle/make_golden.py -> tests/golden/loss_*.npz.
"""
from __future__ import annotations

import numpy as np


def _log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


def mse_logits(S, Tt):
    """kd_loss(KD_logit): 0.5 * MSELoss(reduction='mean')(S, Tt).  -> (loss, dS)"""
    d = S - Tt
    return 0.5 * (d * d).mean(), d / d.size


def dkd(S, Tt, pos, alpha, beta, T):
    """dkd_loss (PROD/ProD_KD/model/models.py:1070-1094): with ps = softmax(S/T), pt = softmax(Tt/T), the binary
    (target, others) KL `TCKD` plus the KL between the distributions renormalised over the non-target columns `NCKD`
    (the reference gets the latter by subtracting 1000 from the target logit: exp(-1000) is exactly 0 in fp32), both
    summed over rows, times T^2 / rows.  -> (alpha TCKD + beta NCKD, dS)"""
    n = S.shape[0]
    rows = np.arange(n)
    ps, pt = np.exp(_log_softmax(S / T)), np.exp(_log_softmax(Tt / T))
    gt = np.zeros_like(S, bool)
    gt[rows, pos] = True
    ps_t, pt_t = ps[gt], pt[gt]
    ps_o, pt_o = (ps * ~gt).sum(1), (pt * ~gt).sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        tck = np.where(pt_t > 0, pt_t * (np.log(pt_t) - np.log(ps_t)), 0.0) + \
              np.where(pt_o > 0, pt_o * (np.log(pt_o) - np.log(ps_o)), 0.0)
        ps2, pt2 = ps / ps_o[:, None], pt / pt_o[:, None]
        nck = np.where(~gt & (pt2 > 0), pt2 * (np.log(pt2) - np.log(ps2)), 0.0).sum(1)
    loss = (alpha * tck.sum() + beta * nck.sum()) * T * T / n
    g_t = np.where(gt, (ps_t - pt_t)[:, None], ps * (pt_t - pt_o * ps_t / ps_o)[:, None])
    g_n = np.where(gt, 0.0, ps2 - pt2)
    return loss, (alpha * g_t + beta * g_n) * T / n


def embed_mse(a, b):
    """0.5 * MSELoss(mean)(a, b) on raw embeddings (one half of bi_logit).  -> (loss, d/da)"""
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return 0.5 * (d * d).mean(), d / d.size


def score_loss(Q, C, *, terms, q_begin=0, q_count=None, c_begin=0, group=1, pos_idx=None, teacher=None,
               w=None, temperature=1.0, eps=1e-7, score_scale=1.0, grad_scale=1.0, dkd_alpha=1.0, dkd_beta=1.0):
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

    local_terms = terms & {"kl_local", "ce_local", "kd_local", "nll_local", "adv_local", "mse_local", "dkd_local"}
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
        if "adv_local" in terms:
            # SimANS/wiki/co_training_wiki_train.py:212-225: reward = log(softmax([t_0, t_j])_0 + eps) on raw logits,
            # adv_loss = sum(reward * log(p_s + eps))  (a sum over the batch, not a mean)
            tr = np.asarray(teacher, np.float64)
            reward = np.log(1.0 / (1.0 + np.exp(tr - tr[:, :1])) + eps)
            loss = (reward * np.log(p + eps)).sum()
            out["adv_local"] = loss
            out["total"] += w.get("adv_local", 1.0) * loss
            rho = reward * p / (p + eps)
            ds += w.get("adv_local", 1.0) * (rho - p * rho.sum(-1, keepdims=True))
        if "mse_local" in terms:
            loss, g = mse_logits(s, np.asarray(teacher, np.float64))
            out["mse_local"] = loss
            out["total"] += w.get("mse_local", 1.0) * loss
            ds += w.get("mse_local", 1.0) * g
        if "dkd_local" in terms:
            loss, g = dkd(s, np.asarray(teacher, np.float64), np.zeros(B, np.int64), dkd_alpha, dkd_beta, temperature)
            out["dkd_local"] = loss
            out["total"] += w.get("dkd_local", 1.0) * loss
            ds += w.get("dkd_local", 1.0) * g
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


def pair_distill(Qa, Ca, Qb, Cb, *, pos_idx, w_nll_a=0.0, w_nll_b=0.0, w_kl_b2a=0.0, w_kl_a2b=0.0, w_kd_b2a=0.0,
                 temperature=1.0, q_begin=0, q_count=None, c_begin=0, c_count=None, grad_scale=1.0, w_mse_b2a=0.0,
                 w_dkd_b2a=0.0, dkd_alpha=1.0, dkd_beta=1.0):
    """LEAD mutual distillation between a student (a) and a teacher (b) dual encoder on the global score matrices
    (/root/reference/LEAD/util.py:594-644 scores, :737-760 assembly; no .detach(): targets carry gradient).
      total = w_nll_a NLL(A) + w_nll_b NLL(B) + w_kl_b2a KL(softmax(B/T)||softmax(A)) + w_kl_a2b KL(softmax(A/T)||softmax(B))
              + w_kd_b2a T^2 KL(softmax(B/T)||softmax(A/T))      (PROD BiEncoderKDLoss KD_softmax, models.py:1050-1056)
              + w_mse_b2a 0.5 MSE(A, B) + w_dkd_b2a DKD(A | B)   (PROD KD_logit / DKD; their teacher is run under
                no_grad, PROD/ProD_KD/run_progressive_distill_marco.py: B is a constant for these two terms)
    Returns (losses dict, dQa, dCa, dQb, dCb) restricted to the local row / column ranges."""
    Qa, Ca, Qb, Cb = (np.asarray(x, np.float64) for x in (Qa, Ca, Qb, Cb))
    GQ, GC = Qa.shape[0], Ca.shape[0]
    q_count = GQ if q_count is None else q_count
    c_count = GC if c_count is None else c_count
    pos = np.asarray(pos_idx, np.int64)
    A, B = Qa @ Ca.T, Qb @ Cb.T
    T = temperature
    lsa, lsb = _log_softmax(A), _log_softmax(B)
    lta, ltb = _log_softmax(A / T), _log_softmax(B / T)
    ta, tb = np.exp(lta), np.exp(ltb)
    rows = np.arange(GQ)
    nll_a, nll_b = -lsa[rows, pos].mean(), -lsb[rows, pos].mean()
    k_b2a_row = (tb * (ltb - lsa)).sum(1)
    k_a2b_row = (ta * (lta - lsb)).sum(1)
    kl_b2a, kl_a2b = k_b2a_row.mean(), k_a2b_row.mean()
    k_kd_row = (tb * (ltb - lta)).sum(1)
    kd_b2a = T * T * k_kd_row.mean()
    total = w_nll_a * nll_a + w_nll_b * nll_b + w_kl_b2a * kl_b2a + w_kl_a2b * kl_a2b + w_kd_b2a * kd_b2a
    oh = np.zeros_like(A)
    oh[rows, pos] = 1.0
    pa, pb = np.exp(lsa), np.exp(lsb)
    dA = (w_nll_a * (pa - oh) + w_kl_b2a * (pa - tb) + w_kl_a2b / T * ta * ((lta - lsb) - k_a2b_row[:, None])
          + w_kd_b2a * T * (ta - tb)) / GQ
    dB = (w_nll_b * (pb - oh) + w_kl_a2b * (pb - ta) + w_kl_b2a / T * tb * ((ltb - lsa) - k_b2a_row[:, None])
          + w_kd_b2a * T * tb * ((ltb - lta) - k_kd_row[:, None])) / GQ
    mse_b2a = dkd_b2a = 0.0
    if w_mse_b2a:
        mse_b2a, g = mse_logits(A, B)
        total, dA = total + w_mse_b2a * mse_b2a, dA + w_mse_b2a * g
    if w_dkd_b2a:
        dkd_b2a, g = dkd(A, B, pos, dkd_alpha, dkd_beta, T)
        total, dA = total + w_dkd_b2a * dkd_b2a, dA + w_dkd_b2a * g
    qs, cs = slice(q_begin, q_begin + q_count), slice(c_begin, c_begin + c_count)
    out = dict(total=total, nll_a=nll_a, nll_b=nll_b, kl_b2a=kl_b2a, kl_a2b=kl_a2b, kd_b2a=kd_b2a,
               n_correct=float((A.argmax(1) == pos).sum()), mse_b2a=mse_b2a, dkd_b2a=dkd_b2a)
    g = grad_scale
    return out, g * dA[qs] @ Ca, g * dA[:, cs].T @ Qa, g * dB[qs] @ Cb, g * dB[:, cs].T @ Qb


def lead_de_db_loss(q_db, c_db, q_de, c_de, qL_db, cL_db, qL_de, cL_de, *, group, temperature=1.0, layer_temperature=1.0,
                    para_db=1.0, para_de=1.0, para_dis=1.0, para_layer=1.0, reweight=True, train_de=True):
    """The DE -> DB branch of LEAD's distill_loss incl. layer_score_dis_loss (/root/reference/LEAD/util.py:539-559,
    594-644, 737-770) on already-gathered embeddings.  qL_*/cL_*: lists of per-selected-layer [CLS] matrices.
    Returns (total, dict of gradients w.r.t. every input)."""
    GQ = np.asarray(q_db).shape[0]
    pos = np.arange(GQ) * group
    out, dqs, dcs, dqt, dct = pair_distill(q_db, c_db, q_de, c_de, pos_idx=pos, w_nll_a=para_db,
                                           w_nll_b=para_de if train_de else 0.0, w_kl_b2a=para_dis,
                                           w_kl_a2b=para_dis if train_de else 0.0, temperature=temperature)
    total = out["total"]
    grads = dict(q_db=dqs, c_db=dcs, q_de=dqt, c_de=dct, qL_db=[], cL_db=[], qL_de=[], cL_de=[])
    layer = [pair_distill(a, b, c, d, pos_idx=pos, w_nll_b=0.0, w_kl_b2a=1.0, temperature=temperature)
             for a, b, c, d in zip(qL_db, cL_db, qL_de, cL_de)]
    if layer:
        nll_t = np.array([l[0]["nll_b"] for l in layer])
        if reweight:  # lambda_i = softmax_i(-NLL(S_de^l) / T_layer), detached (rebuilt via torch.tensor in the reference)
            z = -nll_t / layer_temperature
            lam = np.exp(z - z.max())
            lam /= lam.sum()
        else:
            lam = np.full(len(layer), para_layer)
        for li, l in zip(lam, layer):
            total += li * l[0]["kl_b2a"]
            grads["qL_db"].append(li * l[1]); grads["cL_db"].append(li * l[2])
            grads["qL_de"].append(li * l[3]); grads["cL_de"].append(li * l[4])
    return total, grads


def local_pair_distill(Q, C, teacher, *, group, w_nll_a=0.0, w_nll_b=0.0, w_kl_b2a=0.0, w_kl_a2b=0.0, temperature=1.0):
    """LEAD --distill_ce variant (/root/reference/LEAD/util.py:669-717, 801-829): pair_distill's four terms on each query's
    own [1+n] block; a = q_i . c_ij, b = the cross-encoder's relevance logits, positive = column 0.
    Returns (losses, dQ, dC, d_teacher)."""
    Q, C, Bm = np.asarray(Q, np.float64), np.asarray(C, np.float64), np.asarray(teacher, np.float64)
    nq, H = Q.shape
    Cl = C.reshape(nq, group, H)
    A = np.einsum("bh,bdh->bd", Q, Cl)
    T = temperature
    lsa, lsb = _log_softmax(A), _log_softmax(Bm)
    lta, ltb = _log_softmax(A / T), _log_softmax(Bm / T)
    ta, tb = np.exp(lta), np.exp(ltb)
    nll_a, nll_b = -lsa[:, 0].mean(), -lsb[:, 0].mean()
    kb = (tb * (ltb - lsa)).sum(1)
    ka = (ta * (lta - lsb)).sum(1)
    total = w_nll_a * nll_a + w_nll_b * nll_b + w_kl_b2a * kb.mean() + w_kl_a2b * ka.mean()
    oh = np.zeros_like(A)
    oh[:, 0] = 1.0
    pa, pb = np.exp(lsa), np.exp(lsb)
    dA = (w_nll_a * (pa - oh) + w_kl_b2a * (pa - tb) + w_kl_a2b / T * ta * ((lta - lsb) - ka[:, None])) / nq
    dB = (w_nll_b * (pb - oh) + w_kl_a2b * (pb - ta) + w_kl_b2a / T * tb * ((ltb - lsa) - kb[:, None])) / nq
    out = dict(total=total, nll_a=nll_a, nll_b=nll_b, kl_b2a=kb.mean(), kl_a2b=ka.mean())
    return out, np.einsum("bd,bdh->bh", dA, Cl), (dA[:, :, None] * Q[:, None, :]).reshape(nq * group, H), dB


def lead_ce_db_loss(q_db, c_db, ce_logits, qL_db, cL_db, ce_layer_logits, *, group, temperature=1.0, layer_temperature=1.0,
                    para_db=1.0, para_ce=1.0, para_dis=1.0, train_ce=True):
    """The CE -> DB branch of LEAD's distill_loss with lambda re-weighted per-layer score distillation
    (/root/reference/LEAD/util.py:669-717, 801-829).  ce_layer_logits: list of [B, 1+n] per-selected-layer CE logits.
    Returns (total, grads dict)."""
    out, dq, dc, dt = local_pair_distill(q_db, c_db, ce_logits, group=group, w_nll_a=para_db,
                                         w_nll_b=para_ce if train_ce else 0.0, w_kl_b2a=para_dis,
                                         w_kl_a2b=para_dis if train_ce else 0.0, temperature=temperature)
    total = out["total"]
    grads = dict(q_db=dq, c_db=dc, ce_logits=dt, qL_db=[], cL_db=[], ce_layers=[])
    layer = [local_pair_distill(a, b, t, group=group, w_kl_b2a=1.0, temperature=temperature)
             for a, b, t in zip(qL_db, cL_db, ce_layer_logits)]
    if layer:
        z = -np.array([l[0]["nll_b"] for l in layer]) / layer_temperature
        lam = np.exp(z - z.max())
        lam /= lam.sum()
        for li, l in zip(lam, layer):
            total += li * l[0]["kl_b2a"]
            grads["qL_db"].append(li * l[1]); grads["cL_db"].append(li * l[2]); grads["ce_layers"].append(li * l[3])
    return total, grads
