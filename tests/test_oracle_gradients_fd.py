"""CPU: the oracle's closed-form gradients (oracle/losses_numpy.py) against central finite differences of its own loss
values, for every loss term and on rank-local slices — independent of the golden fixtures (which pin one shape each)."""
import numpy as np
import pytest

from oracle import losses_numpy as ln


def fd_grad(f, x, eps=1e-6):
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        old = x[i]
        x[i] = old + eps
        hi = f()
        x[i] = old - eps
        lo = f()
        x[i] = old
        g[i] = (hi - lo) / (2 * eps)
    return g


LOCAL_TERMS = ["kl_local", "ce_local", "kd_local", "nll_local", "adv_local", "mse_local", "dkd_local"]


@pytest.mark.parametrize("terms", [[t] for t in LOCAL_TERMS] + [["nll_global"], ["nll_global", "kl_local"], LOCAL_TERMS])
def test_score_loss_gradients(terms):
    rng = np.random.RandomState(len(terms) * 7 + len(terms[0]))
    GQ, group, H = 4, 3, 5
    Q, C = 0.7 * rng.standard_normal((GQ, H)), 0.7 * rng.standard_normal((GQ * group, H))
    teacher = 2 * rng.standard_normal((2, group))
    kw = dict(terms=terms, q_begin=1, q_count=2, c_begin=3, group=group, pos_idx=np.arange(GQ) * group, teacher=teacher,
              w={t: 0.3 + 0.2 * i for i, t in enumerate(terms)}, temperature=1.7, score_scale=0.9, dkd_alpha=0.8,
              dkd_beta=2.5)
    out, dQ, dC = ln.score_loss(Q, C, **kw)
    f = lambda: ln.score_loss(Q, C, **kw)[0]["total"]
    gQ, gC = fd_grad(f, Q), fd_grad(f, C)
    # gradients reach the local rows only; everything else is a constant of the rank (LEAD/util.py:200-204 semantics)
    # — but the GLOBAL term's true derivative w.r.t. remote rows is not zero: compare on the local slices
    assert np.abs(gQ[1:3] - dQ).max() < 1e-6
    assert np.abs(gC[3:9] - dC).max() < 1e-6


@pytest.mark.parametrize("weights", [dict(w_nll_a=1.0), dict(w_nll_b=1.0), dict(w_kl_b2a=1.0), dict(w_kl_a2b=1.0),
                                     dict(w_kd_b2a=1.0), dict(w_nll_a=0.3, w_nll_b=0.4, w_kl_b2a=0.5, w_kl_a2b=0.6, w_kd_b2a=0.7)])
def test_pair_distill_gradients_of_both_models(weights):
    rng = np.random.RandomState(3)
    GQ, group, H = 3, 2, 4
    Qa, Ca, Qb, Cb = (0.8 * rng.standard_normal(s) for s in ((GQ, H), (GQ * group, H), (GQ, H), (GQ * group, H)))
    kw = dict(pos_idx=np.arange(GQ) * group, temperature=1.5, **weights)
    out, dQa, dCa, dQb, dCb = ln.pair_distill(Qa, Ca, Qb, Cb, **kw)
    f = lambda: ln.pair_distill(Qa, Ca, Qb, Cb, **kw)[0]["total"]
    for x, g in ((Qa, dQa), (Ca, dCa), (Qb, dQb), (Cb, dCb)):  # no .detach() in LEAD's distill_loss: targets carry gradient
        assert np.abs(fd_grad(f, x) - g).max() < 1e-6


@pytest.mark.parametrize("weights", [dict(w_mse_b2a=1.0), dict(w_dkd_b2a=1.0, dkd_alpha=0.7, dkd_beta=3.0)])
def test_pair_distill_prod_terms_student_gradient(weights):
    rng = np.random.RandomState(4)
    GQ, group, H = 3, 3, 4
    Qa, Ca, Qb, Cb = (0.8 * rng.standard_normal(s) for s in ((GQ, H), (GQ * group, H), (GQ, H), (GQ * group, H)))
    kw = dict(pos_idx=np.arange(GQ) * group, temperature=2.0, **weights)
    out, dQa, dCa, _, _ = ln.pair_distill(Qa, Ca, Qb, Cb, **kw)
    f = lambda: ln.pair_distill(Qa, Ca, Qb, Cb, **kw)[0]["total"]
    assert np.abs(fd_grad(f, Qa) - dQa).max() < 1e-6 and np.abs(fd_grad(f, Ca) - dCa).max() < 1e-6


def test_local_pair_and_embed_mse_gradients():
    rng = np.random.RandomState(5)
    nq, group, H = 3, 4, 5
    Q, C, t = 0.6 * rng.standard_normal((nq, H)), 0.6 * rng.standard_normal((nq * group, H)), rng.standard_normal((nq, group))
    kw = dict(group=group, w_nll_a=0.4, w_nll_b=0.3, w_kl_b2a=0.8, w_kl_a2b=0.6, temperature=1.3)
    res = ln.local_pair_distill(Q, C, t, **kw)
    f = lambda: ln.local_pair_distill(Q, C, t, **kw)[0]["total"]
    grads = res[1:]
    for x, g in zip((Q, C, t), grads):
        assert np.abs(fd_grad(f, x) - g).max() < 1e-6
    a, b = rng.standard_normal((3, 4)), rng.standard_normal((3, 4))
    loss, g = ln.embed_mse(a, b)
    assert np.abs(fd_grad(lambda: ln.embed_mse(a, b)[0], a) - g).max() < 1e-7
