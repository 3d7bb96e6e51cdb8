"""CPU: LEAD's ColBERT-teacher distillation arithmetic (losses.colbert_score, losses.lead_col_distill_loss) against the
reference's own `distill_loss` / `colbert_score` run on the same tensors (tests/golden/lead_col_distill.npz, written by
oracle/make_golden.py from /root/reference/LEAD/util.py:459-465, :561-664, :730-735, :772-794).  These are plain tensor
ops on [GQ, GC]-sized score matrices, so the pin runs without a GPU; the tower that feeds them (HFColBertEncoder on the
CUDA kernels) is checked in tests/test_gpu_hidden_states.py."""
import numpy as np
import pytest
import torch

from simxns_b200.losses import colbert_score, lead_col_distill_loss


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "lead_col_distill.npz")


@pytest.mark.parametrize("tag", ["both", "frozen"])
def test_col_branch_loss_and_every_gradient_vs_reference(gold, tag):
    names = ("q_db", "c_db", "q_col", "c_col", "qL_db", "cL_db", "qL_col", "cL_col")
    t = {n: torch.from_numpy(gold[n].copy()).requires_grad_(True) for n in names}
    train_col, reweight, T, layer_T, p_db, p_col, p_dis, p_layer = [float(x) for x in gold[tag + "_cfg"]]
    loss, parts = lead_col_distill_loss(
        t["q_db"], t["c_db"], t["q_col"], t["c_col"], torch.from_numpy(gold["doc_mask"]), t["qL_db"][:, :, 0], t["cL_db"][:, :, 0],
        t["qL_col"], t["cL_col"], selected_db=gold["sel_db"].tolist(), selected_col=gold["sel_col"].tolist(), temperature=T,
        layer_temperature=layer_T, para_db=p_db, para_col=p_col, para_dis=p_dis, para_layer=p_layer, reweight=bool(reweight),
        train_col=bool(train_col), gather=False)
    assert abs(loss.item() - float(gold[tag + "_loss"][0])) < 1e-5
    loss.backward()
    for n in names:
        want = gold[f"{tag}_d_{n}"]
        got = t[n].grad.numpy() if t[n].grad is not None else np.zeros_like(want)
        assert np.abs(got - want).max() <= 1e-6 + 1e-4 * np.abs(want).max(), n
    assert set(parts) >= {"loss_db", "loss_col_db_dis", "loss_db_col_dis", "loss_col_db_layer_dis"}


def test_colbert_score_chunking_and_masked_documents(gold):
    q = torch.from_numpy(gold["q_col"].copy()).requires_grad_(True)
    d = torch.from_numpy(gold["c_col"].copy()).requires_grad_(True)
    mask = torch.from_numpy(gold["doc_mask"].copy())
    mask[2] = 0  # a document with no kept token: every pair scores -9e9 per query position and passes no gradient
    whole = colbert_score(q, d, mask, q.shape[0], chunk=1 << 20)
    g = torch.randn(whole.shape, generator=torch.Generator().manual_seed(3))
    (whole * g).sum().backward()
    gq, gd = q.grad.clone(), d.grad.clone()
    q.grad = d.grad = None
    parts = colbert_score(q, d, mask, q.shape[0], chunk=5)     # 12 documents in chunks of 5, 5, 2
    (parts * g).sum().backward()
    assert torch.allclose(whole, parts, rtol=1e-6, atol=1e-6)   # BLAS blocks the chunked products differently
    assert torch.allclose(q.grad, gq, atol=1e-6) and torch.allclose(d.grad, gd, atol=1e-6)
    assert torch.all(whole[:, 2] == -9e9 * q.shape[1]) and torch.all(d.grad[2] == 0)
    # padded document tokens never receive gradient
    assert torch.all(d.grad[mask == 0] == 0)
    # the straightforward formulation (util.py:459-465) on the same inputs
    with torch.no_grad():
        logits = torch.einsum("qih,djh->qidj", q, d).masked_fill(mask[None, None] == 0, -9e9)
        assert torch.allclose(whole, logits.max(3).values.sum(1), atol=1e-5)


def test_virt_attention_maps_and_loss_vs_reference(golden_dir):
    """LEAD/util.py:504-537 run by oracle/make_golden.py (tests/golden/lead_virt.npz)"""
    import types

    from simxns_b200.losses import query_doc_attention_map, virt_loss

    g = np.load(golden_dir / "lead_virt.npz")
    t = {n: torch.from_numpy(g[n].copy()).requires_grad_(True) for n in ("q_t", "d_t", "q_s", "d_s")}
    args = types.SimpleNamespace(device="cpu", temperature=2.0)
    ql, dl = g["query_len"].tolist(), g["doc_len"].tolist()
    maps_t, _, tgt_t = query_doc_attention_map(args, t["q_t"], t["d_t"], ql, dl, g["sel_t"].tolist())
    maps_s, lsm_s, _ = query_doc_attention_map(args, t["q_s"], t["d_s"], ql, dl, g["sel_s"].tolist())
    assert np.allclose(maps_t[0].detach().numpy(), g["map_t0"], atol=1e-6)
    assert np.allclose(tgt_t[1].detach().numpy(), g["tgt_t1"], atol=1e-6)
    assert np.allclose(lsm_s[1].detach().numpy(), g["lsm_s1"], rtol=1e-5, atol=1e-3)   # entries are ~ -1e9 where masked
    loss = virt_loss(maps_t, maps_s, 0.7)
    assert abs(loss.item() - float(g["loss"][0])) < 1e-5
    loss.backward()
    for n in t:
        want = g["d_" + n]
        assert np.abs(t[n].grad.numpy() - want).max() <= 1e-7 + 1e-4 * np.abs(want).max(), n


def test_warm_up_losses_vs_reference(gold):
    """get_loss_dual's late-interaction branch and get_loss_cross (LEAD/util.py:410-451), reference values in the fixture"""
    import types

    from simxns_b200.losses import get_loss_cross, get_loss_dual

    args = types.SimpleNamespace(device="cpu", local_rank=-1, model_type="colbert")
    q = torch.from_numpy(gold["q_col"].copy()).requires_grad_(True)
    c = torch.from_numpy(gold["c_col"].copy()).requires_grad_(True)
    loss = get_loss_dual(args, q, c, torch.from_numpy(gold["doc_mask"]))
    assert abs(loss.item() - float(gold["warm_loss"][0])) < 1e-6
    loss.backward()
    assert np.allclose(q.grad.numpy(), gold["warm_d_q_col"], atol=1e-7, rtol=1e-4)
    assert np.allclose(c.grad.numpy(), gold["warm_d_c_col"], atol=1e-7, rtol=1e-4)
    t = torch.from_numpy(gold["ce_logits"].copy()).requires_grad_(True)
    lc = get_loss_cross(args, t)
    lc.backward()
    assert abs(lc.item() - float(gold["ce_warm_loss"][0])) < 1e-6 and np.allclose(t.grad.numpy(), gold["ce_warm_d"], atol=1e-7)
    with pytest.raises(RuntimeError, match="CUDA"):      # the dual-encoder branch is the fused kernel: no CPU path
        get_loss_dual(types.SimpleNamespace(model_type="dual_encoder"), q[:, 0], c[:, 0], None)


def test_colbert_score_random_shapes_vs_direct_formula():
    """seeded sweep over (queries, documents, lengths, width, chunk): forward and both gradients against the direct
    einsum / max / sum formulation under autograd"""
    rng = np.random.RandomState(11)
    for _ in range(12):
        Q, D, Sq, Sd, H = (int(rng.randint(1, 5)), int(rng.randint(1, 9)), int(rng.randint(1, 6)), int(rng.randint(1, 9)),
                           int(rng.choice([4, 16, 32])))
        chunk = int(rng.randint(1, D + 2))
        q = torch.from_numpy(rng.standard_normal((Q, Sq, H)).astype(np.float32)).requires_grad_(True)
        d = torch.from_numpy(rng.standard_normal((D, Sd, H)).astype(np.float32)).requires_grad_(True)
        mask = torch.from_numpy((rng.rand(D, Sd) < 0.7).astype(np.int64))
        mask[:, 0] = 1                                            # at least one kept token per document
        w = torch.from_numpy(rng.standard_normal((Q, D)).astype(np.float32))
        got = colbert_score(q, d, mask, Q, chunk=chunk)
        (got * w).sum().backward()
        gq, gd = q.grad.clone(), d.grad.clone()
        q.grad = d.grad = None
        logits = torch.einsum("qih,djh->qidj", q, d).masked_fill(mask[None, None] == 0, -9e9)
        want = logits.max(3).values.sum(1)
        (want * w).sum().backward()
        assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
        assert torch.allclose(gq, q.grad, atol=1e-5, rtol=1e-4) and torch.allclose(gd, d.grad, atol=1e-5, rtol=1e-4)


def test_oracle_hidden_state_backward_vs_reference_colbert_tower(golden_dir):
    """pins the oracle's `backward(d_hidden=)` (the checker of simx_encoder_add_cotangent) to the reference itself: the numpy
    tower's per-layer states equal those of LEAD's HFColBertEncoder, and pushing the cotangents of the reference's loss
    (colbert_score on the normalised output and on layer 1) through the oracle reproduces the reference's parameter
    gradients (tests/golden/lead_colbert_tiny.npz)"""
    from oracle import bert_numpy as bn

    g = np.load(golden_dir / "lead_colbert_tiny.npz")
    cfg = bn.bert_config(num_layers=3, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64)
    params = bn.synth_params(cfg, seed=int(g["seed"][0]))
    layer = int(g["layer"][0])
    oq, od = bn.EncoderOracle(cfg, params), bn.EncoderOracle(cfg, params)
    hq = np.stack(oq.forward(g["q_ids"], g["q_mask"])["hidden"], axis=1)
    hd = np.stack(od.forward(g["d_ids"], g["d_mask"])["hidden"], axis=1)
    assert np.abs(hq[:, 1:] - g["q_hid"]).max() < 2e-5 and np.abs(hd[:, 1:] - g["d_hid"]).max() < 2e-5
    tq = torch.from_numpy(hq).requires_grad_(True)
    td = torch.from_numpy(hd).requires_grad_(True)
    mask = torch.from_numpy(g["d_mask"])
    norm = torch.nn.functional.normalize
    scores = colbert_score(norm(tq[:, -1], p=2, dim=2), norm(td[:, -1], p=2, dim=2), mask, 2)
    layer_scores = colbert_score(tq[:, layer + 1], td[:, layer + 1], mask, 2) / 64.0
    assert np.abs(scores.detach().numpy() - g["scores"]).max() < 1e-4
    target = torch.tensor([0, 3])
    lsm = torch.nn.functional.log_softmax
    loss = torch.nn.functional.nll_loss(lsm(scores, dim=-1), target) + 0.5 * torch.nn.functional.nll_loss(lsm(layer_scores, dim=-1), target)
    assert abs(loss.item() - float(g["loss"][0])) < 1e-5
    loss.backward()
    gq, gd = oq.backward(d_hidden=tq.grad.numpy()), od.backward(d_hidden=td.grad.numpy())
    worst = 0.0
    scale = max(np.abs(g[k]).max() for k in g.files if k.startswith("grad/"))
    for k in g.files:
        if k.startswith("grad/"):
            n = k[5:]
            worst = max(worst, np.abs(gq[n] + gd[n] - g[k]).max() / max(np.abs(g[k]).max(), 1e-3 * scale))
    assert worst < 2e-4, worst
