"""B200 (`-m gpu`): every KD_type branch of PROD's loss classes through the reference-shaped wrappers
(simxns_b200.losses.CrossBERTKDLoss / BiEncoderKDLoss over simx_score_loss, simx_pair_distill, simx_embed_mse) against
tests/golden/losses_prod_kd.npz — outputs of PROD's own classes (PROD/ProD_KD/model/models.py:669-781, 970-1094) — and
against the oracle on rank-local slices.  fp32 loss kernels: loss within 1e-4, gradients within 1e-5 of the scale."""
import types

import numpy as np
import pytest
import torch

from oracle import losses_numpy as ln
from simxns_b200 import _lib
from simxns_b200.losses import BiEncoderKDLoss, CrossBERTKDLoss, fused_pair_distill, fused_score_loss

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    _lib.require_device()


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir / "losses_prod_kd.npz")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def args_from(g, kd):
    ce_w, kd_w, T, lwf_w, alpha, beta = [float(x) for x in g["hyper"]]
    return types.SimpleNamespace(KD_type=kd, TEMPERATURE=T, CE_WEIGHT=ce_w, KD_WEIGHT=kd_w, LwF_WEIGHT=lwf_w,
                                 DKD_alpha=alpha, DKD_beta=beta)


def check(g, tag, loss, correct, q, c):
    ref = float(g[tag + "_loss"][0])
    assert abs(loss.item() - ref) < 1e-4 * max(1.0, abs(ref)), (tag, loss.item(), ref)
    assert int(correct) == int(g[tag + "_correct"][0]), tag
    loss.backward()
    for got, key in ((q.grad, "_dQ"), (c.grad, "_dC")):
        want = g[tag + key]
        assert np.abs(got.cpu().numpy() - want).max() < 1e-5 * max(1.0, np.abs(want).max()), (tag, key)


def fresh(g):
    return cu(g["Q"]).requires_grad_(True), cu(g["C"]).requires_grad_(True)


@pytest.mark.parametrize("kd", ["KD_logit", "DKD"])
def test_cross_kd_types_vs_prod_golden(g, kd):
    q, c = fresh(g)
    loss, correct = CrossBERTKDLoss().calc(args_from(g, kd), q, c, cu(g["teacher"]))
    check(g, "cross_" + kd, loss, correct, q, c)


@pytest.mark.parametrize("kd", ["KD_softmax", "KD_logit"])
def test_cross_lwf_vs_prod_golden(g, kd):
    q, c = fresh(g)
    loss, correct = CrossBERTKDLoss().calc(args_from(g, kd), q, c, cu(g["teacher"]), LwF=True, ori_q_vector=cu(g["Qo"]),
                                           ori_ctx_vectors=cu(g["Co"]))
    check(g, "lwf_" + kd, loss, correct, q, c)


@pytest.mark.parametrize("kd", ["KD_logit", "DKD", "Bi_logit"])
def test_bi_kd_types_vs_prod_golden(g, kd):
    q, c = fresh(g)
    loss, correct = BiEncoderKDLoss().calc(args_from(g, kd), q, c, cu(g["Qt"]), cu(g["Ct"]), [int(x) for x in g["pos"]])
    check(g, "bi_" + kd, loss, correct, q, c)


def test_bad_kd_type_and_missing_dkd_args_raise(g):
    q, c = fresh(g)
    with pytest.raises(ValueError, match="KD_type"):
        CrossBERTKDLoss().calc(args_from(g, "nope"), q, c, cu(g["teacher"]))
    a = args_from(g, "DKD")
    a.DKD_alpha = None
    with pytest.raises(ValueError, match="DKD_alpha"):
        CrossBERTKDLoss().calc(a, q, c, cu(g["teacher"]))
    with pytest.raises(ValueError, match="LwF"):
        CrossBERTKDLoss().calc(args_from(g, "DKD"), q, c, cu(g["teacher"]), LwF=True, ori_q_vector=cu(g["Qo"]),
                               ori_ctx_vectors=cu(g["Co"]))


@pytest.mark.parametrize("GQ,group,H,T", [(8, 4, 64, 2.0), (16, 16, 768, 4.0), (5, 1, 32, 1.0), (3, 64, 128, 3.0)])
def test_local_mse_dkd_vs_oracle(GQ, group, H, T):
    rng = np.random.RandomState(GQ * 31 + group)
    Q, Cm = 0.3 * rng.standard_normal((GQ, H)), 0.3 * rng.standard_normal((GQ * group, H))
    teacher = 2 * rng.standard_normal((GQ, group))
    terms = ["mse_local"] + (["dkd_local"] if group > 1 else [])
    w = dict(mse_local=0.7, dkd_local=1.3)
    ref, dQ, dC = ln.score_loss(Q, Cm, terms=terms, group=group, teacher=teacher, w=w, temperature=T, dkd_alpha=0.5,
                                dkd_beta=2.0)
    q, c = cu(Q.astype(np.float32)).requires_grad_(True), cu(Cm.astype(np.float32)).requires_grad_(True)
    loss, detail = fused_score_loss(q, c, terms, teacher=cu(teacher.astype(np.float32)), weights=w, temperature=T,
                                    group_size=group, gather=False, dkd_alpha=0.5, dkd_beta=2.0)
    assert abs(loss.item() - ref["total"]) < 1e-4 * max(1.0, abs(ref["total"]))
    assert abs(detail[8].item() - ref["mse_local"]) < 1e-4 * max(1.0, ref["mse_local"])
    loss.backward()
    assert np.abs(q.grad.cpu().numpy() - dQ).max() < 1e-5 * max(1.0, np.abs(dQ).max())
    assert np.abs(c.grad.cpu().numpy() - dC).max() < 1e-5 * max(1.0, np.abs(dC).max())


def test_global_mse_dkd_local_slices_vs_oracle():
    """rank-local rows/columns of the global matrix (what rank 1 of 2 computes), teacher constant"""
    rng = np.random.RandomState(5)
    GQ, group, H = 8, 4, 64
    Qa, Ca, Qb, Cb = (0.4 * rng.standard_normal(s) for s in ((GQ, H), (GQ * group, H), (GQ, H), (GQ * group, H)))
    pos = np.arange(GQ) * group
    kw = dict(pos_idx=pos, w_nll_a=0.2, w_mse_b2a=0.6, w_dkd_b2a=1.1, dkd_alpha=1.0, dkd_beta=4.0, temperature=2.0)
    ref, dQ, dC, _, _ = ln.pair_distill(Qa, Ca, Qb, Cb, q_begin=4, q_count=4, c_begin=16, c_count=16, **kw)
    lib = _lib.lib()
    import ctypes as C
    t = lambda a: cu(a.astype(np.float32))
    cfg = _lib.PairCfg(GQ=GQ, GC=GQ * group, H=H, q_begin=4, q_count=4, c_begin=16, c_count=16, w_nll_a=0.2, w_nll_b=0.0,
                       w_kl_b2a=0.0, w_kl_a2b=0.0, temperature=2.0, grad_scale=1.0, w_kd_b2a=0.0, w_mse_b2a=0.6,
                       w_dkd_b2a=1.1, dkd_alpha=1.0, dkd_beta=4.0)
    ws = torch.empty(lib.simx_pair_distill_workspace_bytes(GQ, GQ * group, H), dtype=torch.uint8, device="cuda")
    losses = torch.empty(8, device="cuda")
    dq, dc = torch.empty(4, H, device="cuda"), torch.empty(16, H, device="cuda")
    ten = [t(Qa), t(Ca), t(Qb), t(Cb)]
    p = cu(pos.astype(np.int32))
    _lib.check(lib.simx_pair_distill(C.byref(cfg), *[x.data_ptr() for x in ten], p.data_ptr(), losses.data_ptr(),
                                     dq.data_ptr(), dc.data_ptr(), None, None, ws.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream), "simx_pair_distill")
    torch.cuda.synchronize()
    assert abs(losses[0].item() - ref["total"]) < 1e-4 * max(1.0, abs(ref["total"]))
    assert abs(losses[7].item() - (ref["mse_b2a"] + ref["dkd_b2a"])) < 1e-4 * max(1.0, ref["mse_b2a"] + ref["dkd_b2a"])
    assert np.abs(dq.cpu().numpy() - dQ).max() < 1e-5 * max(1.0, np.abs(dQ).max())
    assert np.abs(dc.cpu().numpy() - dC).max() < 1e-5 * max(1.0, np.abs(dC).max())


def test_prod_reranker_binary_head_vs_oracle():
    """PROD's Reranker (models.py:1232-1254): (binary_logits [N,M,2], relevance_logits [N,M], None) from the [CLS] row"""
    from oracle import bert_numpy as bn
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.prod_models import HFBertEncoder, Reranker

    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=64)
    params = bn.synth_params(cfg, seed=8)
    ec = EncoderConfig(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=64)
    enc = HFBertEncoder(ec, precision=_lib.PREC_BF16X3)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    rr = Reranker(enc, 128).cuda()
    assert sorted(k for k in rr.state_dict() if not k.startswith("encoder.")) == [
        "binary.bias", "binary.weight", "qa_classifier.bias", "qa_classifier.weight"]
    ids, mask = bn.synth_tokens(6, 40, cfg["vocab"], seed=2, ragged=True, min_len=5)
    cls = bn.EncoderOracle(cfg, params).forward(ids, mask)["cls"][:, -1]
    b, r, none = rr(cu(ids.reshape(2, 3, 40)), cu(mask.reshape(2, 3, 40)))
    assert none is None and tuple(b.shape) == (2, 3, 2) and tuple(r.shape) == (2, 3)
    wb, bb = rr.binary.weight.detach().cpu().double().numpy(), rr.binary.bias.detach().cpu().double().numpy()
    wq, bq = rr.qa_classifier.weight.detach().cpu().double().numpy(), rr.qa_classifier.bias.detach().cpu().double().numpy()
    assert np.abs(b.detach().cpu().numpy().reshape(6, 2) - (cls @ wb.T + bb)).max() < 1e-4
    assert np.abs(r.detach().cpu().numpy().reshape(6) - (cls @ wq.T + bq)[:, 0]).max() < 1e-4
    (b.sum() + r.sum()).backward()
    assert rr.binary.weight.grad is not None and enc.embeddings.word_embeddings.weight.grad is not None
