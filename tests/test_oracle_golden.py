"""The oracle (oracle/*.py, numpy float64) pinned against fixtures produced by the reference's own code
(oracle/make_golden.py ran /root/reference's classes on CPU fp32).  Tolerances are fp32 round-off of the reference."""
import numpy as np
import pytest

from oracle import bert_numpy as bn
from oracle import losses_numpy as ln


def _cfg_from(arr, arch="bert"):
    L, H, nh, I, V, P, Tt = [int(x) for x in arr]
    return bn.bert_config(num_layers=L, hidden=H, heads=nh, intermediate=I, vocab=V, max_pos=P, type_vocab=Tt, arch=arch)


def _close(a, b, atol, rtol=1e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    err = np.abs(a - b).max()
    assert err <= atol + rtol * np.abs(b).max(), f"max err {err:.3e} (ref max {np.abs(b).max():.3e})"


def test_encoder_tiny_bert_forward_backward(golden_dir):
    g = np.load(golden_dir / "encoder_tiny_bert.npz")
    cfg = _cfg_from(g["cfg"])
    params = bn.synth_params(cfg, seed=int(g["seed"][0]))
    enc = bn.EncoderOracle(cfg, params)
    out = enc.forward(g["ids"], g["mask"])
    _close(out["cls"], g["cls"], 2e-5)
    _close(out["seq"], g["seq"], 2e-5)
    _close(out["cls"][:, -1], g["pooled"], 2e-5)
    grads = enc.backward(d_cls=g["d_cls"].astype(np.float64), d_seq=g["d_seq"].astype(np.float64))
    n_checked = 0
    for name in bn.param_names(cfg):
        ref = g["grad/" + name]
        _close(grads[name], ref, 1e-5 * max(1.0, np.abs(ref).max()), rtol=2e-4)
        n_checked += 1
    assert n_checked == 5 + 16 * 2 + 2
    # pooler is dead math in the reference wrapper (SimANS/model/models.py:80): gradients exist and are zero
    assert np.all(g["grad/pooler.dense.weight"] == 0) and np.all(grads["pooler.dense.weight"] == 0)


def test_encoder_tiny_distilbert(golden_dir):
    g = np.load(golden_dir / "encoder_tiny_distilbert.npz")
    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64, arch="distilbert")
    enc = bn.EncoderOracle(cfg, bn.synth_params(cfg, seed=int(g["seed"][0])))
    out = enc.forward(g["ids"], g["mask"])
    _close(out["seq"], g["seq"], 2e-5)
    # LEAD's all-layer hidden states exclude the embedding output: entry k == our cls[:, k+1]
    assert int(g["n_hidden"][0]) == 2
    _close(out["cls"][:, 1:], g["hidden_cls"], 2e-5)
    grads = enc.backward(d_seq=g["d_seq"].astype(np.float64))
    for name in bn.param_names(cfg):
        ref = g["grad/" + name]
        _close(grads[name], ref, 1e-5 * max(1.0, np.abs(ref).max()), rtol=2e-4)


def test_reranker_tiny(golden_dir):
    """Reranker = encoder + Linear(H,1) on [CLS] (SimANS/model/models.py:638-659), CE on column 0 (teacher step)"""
    g = np.load(golden_dir / "reranker_tiny.npz")
    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64)
    enc = bn.EncoderOracle(cfg, bn.synth_params(cfg, seed=int(g["seed"][0])))
    N, M, L = g["ids"].shape
    out = enc.forward(g["ids"].reshape(N * M, L), g["mask"].reshape(N * M, L))
    cls = out["cls"][:, -1]
    logits = (cls @ g["qa_w"].astype(np.float64).T + g["qa_b"]).reshape(N, M)
    _close(logits, g["logits"], 2e-5)
    lsm = logits - logits.max(1, keepdims=True)
    lsm = lsm - np.log(np.exp(lsm).sum(1, keepdims=True))
    assert abs(-lsm[:, 0].mean() - g["loss"][0]) < 1e-5
    dlogits = np.exp(lsm)
    dlogits[:, 0] -= 1
    dlogits /= N
    d_cls = np.zeros_like(out["cls"])
    d_cls[:, -1] = dlogits.reshape(-1, 1) * g["qa_w"].astype(np.float64)
    grads = enc.backward(d_cls=d_cls)
    _close(dlogits.reshape(1, -1) @ cls, g["grad/qa_classifier.weight"], 1e-6)
    for name in bn.param_names(cfg):
        ref = g["grad/encoder." + name]
        _close(grads[name], ref, 1e-6 * max(1.0, np.abs(ref).max()) + 1e-7, rtol=3e-4)


def test_losses_against_reference(golden_dir):
    g = np.load(golden_dir / "losses.npz")
    Q, C, t, group = g["Q"], g["C"], g["teacher"], int(g["group"][0])
    GQ = Q.shape[0]
    # BiEncoderNllLoss.calc
    out, dQ, dC = ln.score_loss(Q, C, terms=["nll_global"], group=group, pos_idx=g["nll_pos"])
    assert abs(out["total"] - g["nll_loss"][0]) < 1e-5 and out["n_correct"] == g["nll_correct"][0]
    _close(dQ, g["nll_dQ"], 1e-6)
    _close(dC, g["nll_dC"], 1e-6)
    # SimANS KL with eps inside the log
    for T in (1, 2):
        out, dQ, dC = ln.score_loss(Q, C, terms=["kl_local"], group=group, teacher=t, temperature=float(T))
        assert abs(out["total"] - g[f"kl_T{T}_loss"][0]) < 1e-5
        _close(dQ, g[f"kl_T{T}_dQ"], 1e-6)
        _close(dC, g[f"kl_T{T}_dC"], 1e-6)
    # PROD CrossBERTKDLoss: 0.1 * NLL(col 0) + 0.9 * T^2 KL(softmax(t/T) || softmax(s/T)), T = 4
    out, dQ, dC = ln.score_loss(Q, C, terms=["nll_local", "kd_local"], group=group, teacher=t, temperature=4.0,
                                w=dict(nll_local=0.1, kd_local=0.9))
    assert abs(out["total"] - g["prod_loss"][0]) < 1e-5
    _close(dQ, g["prod_dQ"], 1e-6)
    _close(dC, g["prod_dC"], 1e-6)
    # PROD prob_loss / MASTER CE-to-teacher: 0.1 * NLL + 0.9 * CE
    out, dQ, dC = ln.score_loss(Q, C, terms=["nll_local", "ce_local"], group=group, teacher=t, temperature=1.0,
                                w=dict(nll_local=0.1, ce_local=0.9))
    assert abs(out["total"] - g["prob_loss"][0]) < 1e-5
    _close(dQ, g["prob_dQ"], 1e-6)
    _close(dC, g["prob_dC"], 1e-6)


def test_loss_known_answers():
    """closed-form KATs: equal embeddings -> InfoNCE = log(GC); student == teacher -> KL ~ 0;
    KL batchmean divides by B, not B*(1+n)."""
    H, GQ, group = 16, 4, 3
    Q = np.ones((GQ, H))
    C = np.ones((GQ * group, H))
    out, dQ, dC = ln.score_loss(Q, C, terms=["nll_global"], group=group, pos_idx=np.arange(GQ) * group)
    assert abs(out["total"] - np.log(GQ * group)) < 1e-12
    rng = np.random.RandomState(0)
    Q, C = rng.standard_normal((GQ, H)), rng.standard_normal((GQ * group, H))
    s = np.einsum("bh,bdh->bd", Q, C.reshape(GQ, group, H))
    out, _, _ = ln.score_loss(Q, C, terms=["kl_local"], group=group, teacher=s, eps=0.0)
    assert abs(out["total"]) < 1e-12
    t = rng.standard_normal((GQ, group))
    out1, _, _ = ln.score_loss(Q, C, terms=["kl_local"], group=group, teacher=t)
    p = np.exp(s - s.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
    tt = np.exp(t - t.max(1, keepdims=True)); tt /= tt.sum(1, keepdims=True)
    manual = (tt * (np.log(tt) - np.log(p + 1e-7))).sum() / GQ
    assert abs(out1["total"] - manual) < 1e-12


def test_loss_gradients_finite_difference():
    rng = np.random.RandomState(1)
    GQ, group, H = 5, 3, 8
    Q, C = rng.standard_normal((GQ, H)), rng.standard_normal((GQ * group, H))
    t = rng.standard_normal((2, group))
    kw = dict(terms=["nll_global", "kl_local", "ce_local", "kd_local", "nll_local"], q_begin=2, q_count=2, c_begin=6,
              group=group, pos_idx=np.arange(GQ) * group, teacher=t, temperature=2.0, score_scale=0.5,
              w=dict(nll_global=0.2, kl_local=1.0, ce_local=0.3, kd_local=0.9, nll_local=0.1))
    out, dQ, dC = ln.score_loss(Q, C, **kw)
    h = 1e-6
    for (arr, grad, off) in ((Q, dQ, 2), (C, dC, 6)):
        for i in range(grad.shape[0]):
            for j in (0, H - 1):
                arr[off + i, j] += h
                lp = ln.score_loss(Q, C, **kw)[0]["total"]
                arr[off + i, j] -= 2 * h
                lm = ln.score_loss(Q, C, **kw)[0]["total"]
                arr[off + i, j] += h
                assert abs((lp - lm) / (2 * h) - grad[i, j]) < 1e-6


def test_lead_distill_loss_against_reference(golden_dir):
    """oracle lead_de_db_loss vs the real LEAD distill_loss (DE -> DB, per-layer score distillation, lambda re-weighting)"""
    g = np.load(golden_dir / "lead_distill.npz")
    group = int(g["group"][0])
    sel_db, sel_de = list(g["sel_db"]), list(g["sel_de"])
    for tag in ("both", "frozen"):
        train_de, reweight, T, Tl, p_db, p_de, p_dis, p_layer = g[tag + "_cfg"]
        total, grads = ln.lead_de_db_loss(
            g["q_db"], g["c_db"], g["q_de"], g["c_de"],
            [g["qL_db"][:, l, 0] for l in sel_db], [g["cL_db"][:, l, 0] for l in sel_db],
            [g["qL_de"][:, l, 0] for l in sel_de], [g["cL_de"][:, l, 0] for l in sel_de],
            group=group, temperature=T, layer_temperature=Tl, para_db=p_db, para_de=p_de, para_dis=p_dis,
            para_layer=p_layer, reweight=bool(reweight), train_de=bool(train_de))
        assert abs(total - g[tag + "_loss"][0]) < 2e-5
        for n in ("q_db", "c_db", "q_de", "c_de"):
            _close(grads[n], g[f"{tag}_d_{n}"], 2e-6)
        for n, sel in (("qL_db", sel_db), ("cL_db", sel_db), ("qL_de", sel_de), ("cL_de", sel_de)):
            full = g[f"{tag}_d_{n}"]
            for i, l in enumerate(sel):
                _close(grads[n][i], full[:, l, 0], 2e-6)
            rest = [l for l in range(full.shape[1]) if l not in sel]
            assert np.abs(full[:, rest]).max() == 0.0


def test_prod_biencoder_kd_loss_against_reference(golden_dir):
    """PROD BiEncoderKDLoss (global DE -> DE KD_softmax, CE 0.1 / KD 0.9 / T 4)"""
    g = np.load(golden_dir / "losses.npz")
    out, dQ, dC, _, _ = ln.pair_distill(g["Q"], g["C"], g["Qt"], g["Ct"], pos_idx=g["nll_pos"], w_nll_a=0.1, w_kd_b2a=0.9,
                                        temperature=4.0)
    assert abs(out["total"] - g["bikd_loss"][0]) < 1e-5 and out["n_correct"] == g["bikd_correct"][0]
    _close(dQ, g["bikd_dQ"], 1e-6)
    _close(dC, g["bikd_dC"], 1e-6)


def test_wiki_adversarial_loss_against_reference(golden_dir):
    """SimANS NQ/TQ retriever loss: adv_lambda * sum(reward log p_s) + (1 - adv_lambda) * CE-to-teacher, T_normal = 2"""
    g = np.load(golden_dir / "losses.npz")
    group = int(g["group"][0])
    for lam, tag in ((0.0, "wiki_l0"), (0.3, "wiki_l3")):
        terms = ["ce_local"] + (["adv_local"] if lam else [])
        out, dQ, dC = ln.score_loss(g["Q"], g["C"], terms=terms, group=group, teacher=g["teacher"], temperature=2.0,
                                    w=dict(ce_local=1 - lam, adv_local=lam))
        assert abs(out["total"] - g[tag + "_loss"][0]) < 2e-5
        _close(dQ, g[tag + "_dQ"], 2e-6)
        _close(dC, g[tag + "_dC"], 2e-6)


def test_lead_ce_teacher_distill_against_reference(golden_dir):
    """oracle lead_ce_db_loss vs the real LEAD distill_loss with --distill_ce --train_ce and per-layer CE logits"""
    g = np.load(golden_dir / "lead_distill.npz")
    group = int(g["group"][0])
    sel_db, sel_t = list(g["sel_db"]), list(g["sel_de"])
    T, Tl, p_db, p_ce, p_dis = g["ce_cfg"]
    total, grads = ln.lead_ce_db_loss(g["q_db"], g["c_db"], g["ce_logits"], [g["qL_db"][:, l, 0] for l in sel_db],
                                      [g["cL_db"][:, l, 0] for l in sel_db], [g["ce_layers"][l] for l in sel_t], group=group,
                                      temperature=T, layer_temperature=Tl, para_db=p_db, para_ce=p_ce, para_dis=p_dis)
    assert abs(total - g["ce_loss"][0]) < 2e-5
    _close(grads["q_db"], g["ce_d_q_db"], 2e-6)
    _close(grads["c_db"], g["ce_d_c_db"], 2e-6)
    _close(grads["ce_logits"], g["ce_d_ce_logits"], 2e-6)
    for i, (ls, lt) in enumerate(zip(sel_db, sel_t)):
        _close(grads["qL_db"][i], g["ce_d_qL_db"][:, ls, 0], 2e-6)
        _close(grads["cL_db"][i], g["ce_d_cL_db"][:, ls, 0], 2e-6)
        _close(grads["ce_layers"][i], g["ce_d_ce_layers"][lt], 2e-6)
