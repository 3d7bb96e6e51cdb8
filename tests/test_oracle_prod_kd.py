"""CPU: the oracle's KD_logit / DKD / Bi_logit / LwF restatements against tests/golden/losses_prod_kd.npz, produced by
PROD's own CrossBERTKDLoss / BiEncoderKDLoss (oracle/make_golden.py prod_kd; PROD/ProD_KD/model/models.py:669-781,
970-1094).  fp32 round-off of the reference bounds the tolerances."""
import numpy as np
import pytest

from oracle import losses_numpy as ln


def _close(a, b, tol):
    err = np.abs(np.asarray(a, np.float64) - b).max()
    assert err <= tol * max(1.0, np.abs(b).max()), err


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir / "losses_prod_kd.npz")


def oracle_cross(g, kd, lwf=False):
    ce_w, kd_w, T, lwf_w, alpha, beta = [float(x) for x in g["hyper"]]
    group = int(g["group"][0])
    term = {"KD_logit": "mse_local", "DKD": "dkd_local", "KD_softmax": "kd_local"}[kd]
    out, dQ, dC = ln.score_loss(g["Q"], g["C"], terms=["nll_local", term], group=group, teacher=g["teacher"],
                                w={"nll_local": ce_w, term: kd_w}, temperature=T, dkd_alpha=alpha, dkd_beta=beta)
    total = out["total"]
    if lwf:  # LwF_WEIGHT * kd_loss(scores, scores of the frozen initial student)   (models.py:748-750)
        ori = np.einsum("bh,bdh->bd", g["Qo"].astype(np.float64), g["Co"].astype(np.float64).reshape(-1, group, g["Q"].shape[1]))
        o2, dQ2, dC2 = ln.score_loss(g["Q"], g["C"], terms=[term], group=group, teacher=ori, w={term: lwf_w}, temperature=T)
        total, dQ, dC = total + o2["total"], dQ + dQ2, dC + dC2
    return total, dQ, dC


@pytest.mark.parametrize("kd,lwf", [("KD_logit", False), ("DKD", False), ("KD_softmax", True), ("KD_logit", True)])
def test_cross_kd_variants(g, kd, lwf):
    tag = ("lwf_" if lwf else "cross_") + kd
    total, dQ, dC = oracle_cross(g, kd, lwf)
    assert abs(total - g[tag + "_loss"][0]) < 2e-5 * max(1.0, abs(total))
    _close(dQ, g[tag + "_dQ"], 2e-5)
    _close(dC, g[tag + "_dC"], 2e-5)


@pytest.mark.parametrize("kd", ["KD_logit", "DKD"])
def test_bi_kd_variants(g, kd):
    ce_w, kd_w, T, _, alpha, beta = [float(x) for x in g["hyper"]]
    kw = dict(w_mse_b2a=kd_w) if kd == "KD_logit" else dict(w_dkd_b2a=kd_w, dkd_alpha=alpha, dkd_beta=beta)
    out, dQ, dC, _, _ = ln.pair_distill(g["Q"], g["C"], g["Qt"], g["Ct"], pos_idx=g["pos"], w_nll_a=ce_w, temperature=T, **kw)
    assert abs(out["total"] - g[f"bi_{kd}_loss"][0]) < 2e-5 * max(1.0, abs(out["total"]))
    assert out["n_correct"] == g[f"bi_{kd}_correct"][0]
    _close(dQ, g[f"bi_{kd}_dQ"], 2e-5)
    _close(dC, g[f"bi_{kd}_dC"], 2e-5)


def test_bi_logit(g):
    ce_w, kd_w = float(g["hyper"][0]), float(g["hyper"][1])
    out, dQ, dC = ln.score_loss(g["Q"], g["C"], terms=["nll_global"], group=int(g["group"][0]), pos_idx=g["pos"],
                                w={"nll_global": ce_w})
    lq, gq = ln.embed_mse(g["Q"], g["Qt"])
    lc, gc = ln.embed_mse(g["C"], g["Ct"])
    total = out["total"] + kd_w * (lq + lc)
    assert abs(total - g["bi_Bi_logit_loss"][0]) < 2e-5
    _close(dQ + kd_w * gq, g["bi_Bi_logit_dQ"], 2e-5)
    _close(dC + kd_w * gc, g["bi_Bi_logit_dC"], 2e-5)
