"""Parity tests proper (B200, `-m gpu`): the CUDA path, called through the C ABI (ctypes -> libsimx_b200.so), against
(1) the golden fixtures produced by the reference's own code and (2) the numpy oracle on seeded inputs.

Tolerances (north_star): fp32-class mode (SIMX_PREC_BF16X3) — embeddings/logits and gradients within 1e-3, loss within
1e-4 of the reference (we assert much tighter where the data allows).  bf16 throughput mode is checked at the
deviation the reference itself shows under bf16 autocast (~3 % on logits): 5e-2 relative on hidden
states, 8e-2 of the per-tensor scale on gradients.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import bert_numpy as bn
from oracle import losses_numpy as ln
from simxns_b200 import _lib
from simxns_b200.encoder import EncoderConfig
from simxns_b200.losses import BiEncoderNllLoss, CrossBERTKDLoss, fused_score_loss, simans_distill_loss
from simxns_b200.models import BiBertEncoder, HFBertEncoder, HFDistilBertEncoder, Reranker

pytestmark = pytest.mark.gpu

X3, BF16 = _lib.PREC_BF16X3, _lib.PREC_BF16


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    _lib.require_device()  # fails loudly when the library cannot run here: no fallback path exists


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def ec_from(cfg):
    return EncoderConfig(arch=cfg["arch"], num_layers=cfg["num_layers"], hidden=cfg["hidden"], heads=cfg["heads"],
                         intermediate=cfg["intermediate"], vocab=cfg["vocab"], max_pos=cfg["max_pos"],
                         type_vocab=cfg["type_vocab"], ln_eps=cfg["ln_eps"])


def build(cfg, params, precision):
    cls = HFBertEncoder if cfg["arch"] == "bert" else HFDistilBertEncoder
    m = cls(ec_from(cfg), precision=precision)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    return m.cuda()


def max_rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def assert_grads(model, ref_grads, tol, prefix="", zero_true=()):
    scales = [np.abs(v).max() for v in ref_grads.values() if np.abs(v).max() > 0]
    floor = 1e-3 * float(np.median(scales))
    worst = ("", 0.0, 0.0, 0.0)
    for name, p in model.named_parameters():
        assert p.grad is not None, f"{name} received no gradient (DDP needs dense grads for every parameter)"
        r = ref_grads[prefix + name]
        # tensors whose true gradient is exactly zero (key biases: softmax shift invariance; the reranker head bias
        # under CE: rows of softmax - onehot sum to zero) are judged on the scale of the largest gradient
        is_zero = "key.bias" in name or "k_lin.bias" in name or name in zero_true
        fl = max(scales) if is_zero else floor
        e = float(np.abs(p.grad.cpu().numpy() - r).max() / max(np.abs(r).max(), fl))
        if e > worst[1]:
            worst = (name, e, float(np.abs(r).max()), fl)
    assert worst[1] < tol, (f"gradient of {worst[0]} off by {worst[1]:.3e} (tol {tol}); |ref|max {worst[2]:.3e}, "
                            f"scale floor {worst[3]:.3e}, largest gradient scale {max(scales):.3e}")


# ---------------------------------------------------------------------------------------------------------
# encoder tower vs the reference's own outputs (golden fixtures)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol_h,tol_g", [(X3, 2e-5, 1e-4), (BF16, 5e-2, 8e-2)])
def test_encoder_tiny_bert_vs_reference_golden(golden_dir, precision, tol_h, tol_g):
    g = np.load(golden_dir / "encoder_tiny_bert.npz")
    L, H, nh, I, V, P, Tt = [int(x) for x in g["cfg"]]
    cfg = bn.bert_config(num_layers=L, hidden=H, heads=nh, intermediate=I, vocab=V, max_pos=P, type_vocab=Tt)
    model = build(cfg, bn.synth_params(cfg, seed=int(g["seed"][0])), precision)
    seq, pooled, none = model(input_ids=cu(g["ids"]), attention_mask=cu(g["mask"]))
    assert none is None
    cls = model.last_cls_layers
    valid = g["mask"].astype(bool)
    assert max_rel(seq.detach().cpu().numpy()[valid], g["seq"][valid]) < tol_h
    assert max_rel(pooled.detach().cpu().numpy(), g["pooled"]) < tol_h
    assert max_rel(cls.detach().cpu().numpy(), g["cls"]) < tol_h
    if precision == X3:  # absolute bound of the north_star: 1e-3 on outputs
        assert np.abs(pooled.detach().cpu().numpy() - g["pooled"]).max() < 1e-3
    d_seq = g["d_seq"] * g["mask"][:, :, None]  # the reference's padded rows feed nothing downstream either
    loss = (cls * cu(g["d_cls"])).sum() + (seq * cu(d_seq.astype(np.float32))).sum()
    loss.backward()
    # golden grads were taken with the unmasked d_seq: recompute the reference gradient with the oracle (pinned to the
    # golden in tests/test_oracle_golden.py) for the masked cotangent
    ora = bn.EncoderOracle(cfg, bn.synth_params(cfg, seed=int(g["seed"][0])))
    ora.forward(g["ids"], g["mask"])
    ref = ora.backward(d_cls=g["d_cls"].astype(np.float64), d_seq=d_seq.astype(np.float64))
    assert_grads(model, ref, tol_g)
    assert float(model.pooler.dense.weight.grad.abs().max()) == 0.0  # dead pooler: zero, not None


def test_encoder_full_golden_gradients_fp32_class(golden_dir):
    """every one of the reference's gradients (unmasked random cotangent on all positions), fp32-class mode"""
    g = np.load(golden_dir / "encoder_tiny_bert.npz")
    L, H, nh, I, V, P, Tt = [int(x) for x in g["cfg"]]
    cfg = bn.bert_config(num_layers=L, hidden=H, heads=nh, intermediate=I, vocab=V, max_pos=P, type_vocab=Tt)
    model = build(cfg, bn.synth_params(cfg, seed=int(g["seed"][0])), X3)
    seq, _, _ = model(input_ids=cu(g["ids"]), attention_mask=cu(g["mask"]))
    ((model.last_cls_layers * cu(g["d_cls"])).sum() + (seq * cu(g["d_seq"])).sum()).backward()
    ref = {k[5:]: g[k] for k in g.files if k.startswith("grad/")}
    assert_grads(model, ref, 1e-3)


def test_encoder_tiny_distilbert_vs_lead_golden(golden_dir):
    g = np.load(golden_dir / "encoder_tiny_distilbert.npz")
    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64, arch="distilbert")
    model = build(cfg, bn.synth_params(cfg, seed=int(g["seed"][0])), X3)
    seq, pooled, _ = model(input_ids=cu(g["ids"]), attention_mask=cu(g["mask"]))
    valid = g["mask"].astype(bool)
    assert max_rel(seq.detach().cpu().numpy()[valid], g["seq"][valid]) < 2e-5
    # LEAD's all-layer hidden states (no embedding entry): layer k == cls_layers[:, k+1]
    assert max_rel(model.last_cls_layers[:, 1:].detach().cpu().numpy(), g["hidden_cls"]) < 2e-5
    (seq * cu(g["d_seq"])).sum().backward()
    ref = {k[5:]: g[k] for k in g.files if k.startswith("grad/")}
    assert_grads(model, ref, 1e-3)


def test_reranker_teacher_step_vs_golden(golden_dir):
    """co_training_marco_train.py:225-245 (teacher step): CE over [N, M] relevance logits, column 0 positive"""
    g = np.load(golden_dir / "reranker_tiny.npz")
    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64)
    enc = build(cfg, bn.synth_params(cfg, seed=int(g["seed"][0])), X3)
    rr = Reranker(enc, cfg["hidden"]).cuda()
    rr.qa_classifier.weight.data = cu(g["qa_w"])
    rr.qa_classifier.bias.data = cu(g["qa_b"])
    logits = rr(cu(g["ids"]), cu(g["mask"]))
    assert np.abs(logits.detach().cpu().numpy() - g["logits"]).max() < 1e-4
    loss = torch.nn.CrossEntropyLoss()(logits, torch.zeros(2, dtype=torch.long, device="cuda"))
    assert abs(loss.item() - g["loss"][0]) < 1e-4
    loss.backward()
    ref = {k[5:]: g[k] for k in g.files if k.startswith("grad/")}
    # sum_m (softmax - onehot) = 0 per query: the head bias and the last LayerNorm bias (= W^T sum of dlogits) have an
    # exactly-zero true gradient, the reference's values for them are fp32 round-off
    assert_grads(rr, ref, 1e-3, zero_true=("qa_classifier.bias", "encoder.encoder.layer.1.output.LayerNorm.bias"))


def test_biencoder_cfg1_bert_base_vs_reference_golden(golden_dir):
    """BASELINE.json configs[0]/[1]: BiBertEncoder bert-base, B=2, q 32 / p 128, 1+1 passages, SimANS KL loss:
    embeddings within 1e-3, loss within 1e-4, gradients within 1e-3 of the tensor scale."""
    g = np.load(golden_dir / "biencoder_cfg1.npz")
    cfg = bn.bert_config()
    model = BiBertEncoder(None, question_model=build(cfg, bn.synth_params(cfg, seed=0, perturb=False), X3),
                          ctx_model=build(cfg, bn.synth_params(cfg, seed=1, perturb=False), X3))
    q, a = model(cu(g["q_ids"]), cu(g["q_mask"]), cu(g["p_ids"]), cu(g["p_mask"]))
    assert np.abs(q.detach().cpu().numpy() - g["q_embs"]).max() < 1e-3
    assert np.abs(a.detach().cpu().numpy() - g["a_embs"]).max() < 1e-3
    loss = simans_distill_loss(q, a, cu(g["teacher"]), temperature_distill=1.0)
    assert abs(loss.item() - g["loss"][0]) < 1e-4
    loss.backward()
    n = 0
    scales = [g[k][2] for k in g.files if k.startswith("gstat/") and g[k][2] > 0]
    floor = 1e-3 * float(np.median(scales))
    for name, p in model.named_parameters():
        stat, head = g["gstat/" + name], g["ghead/" + name]
        gr = p.grad.double().reshape(-1)
        # true-zero gradients (key biases; the passage tower's last LayerNorm bias = sum_j dS_ij q_i with sum_j dS_ij = 0
        # for a KL/softmax row): the reference's own values are fp32 round-off, judge them on the largest scale
        zero_true = "key.bias" in name or name == "ctx_model.encoder.layer.11.output.LayerNorm.bias"
        scale = max(scales) if zero_true else max(stat[2], floor)
        assert abs(gr.abs().max().item() - stat[2]) <= 1e-3 * scale + 1e-12, name
        assert np.abs(gr[:16].cpu().numpy() - head).max() <= 1e-3 * scale, name
        if not zero_true:
            assert abs(gr.norm().item() - stat[0]) <= 2e-3 * max(stat[0], floor), name
        if ("grad/" + name) in g.files:
            assert np.abs(p.grad.cpu().numpy() - g["grad/" + name]).max() <= 1e-3 * scale, name
        n += 1
    assert n == 398


# ---------------------------------------------------------------------------------------------------------
# encoder vs oracle: ragged / edge shapes
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,ragged", [(1, 1, False), (2, 7, True), (3, 64, True), (2, 65, True), (5, 128, False),
                                        (2, 160, True), (1, 300, True), (1, 512, True)])
def test_encoder_shapes_vs_oracle(B, S, ragged):
    cfg = bn.bert_config(num_layers=1, hidden=128, heads=2, intermediate=192, vocab=300, max_pos=512)
    params = bn.synth_params(cfg, seed=100 + S)
    ids, mask = bn.synth_tokens(B, S, cfg["vocab"], seed=S, ragged=ragged, min_len=1)
    ora = bn.EncoderOracle(cfg, params)
    out = ora.forward(ids, mask)
    rng = np.random.RandomState(S)
    d_cls = rng.standard_normal(out["cls"].shape)
    ref = ora.backward(d_cls=d_cls)
    for precision, tol_h, tol_g in ((X3, 2e-5, 2e-4), (BF16, 5e-2, 1e-1)):
        model = build(cfg, params, precision)
        cls, _ = model.encode(cu(ids), cu(mask))
        assert max_rel(cls.detach().cpu().numpy(), out["cls"]) < tol_h
        (cls * cu(d_cls, torch.float32)).sum().backward()
        assert_grads(model, ref, tol_g)


def test_encoder_mask_none_and_inference_mode():
    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=64)
    params = bn.synth_params(cfg, seed=3)
    ids, mask = bn.synth_tokens(4, 32, cfg["vocab"], seed=9)
    out = bn.EncoderOracle(cfg, params).forward(ids, mask)
    model = build(cfg, params, X3)
    with torch.no_grad():  # inference path: nothing saved, workspace ping-pong buffers
        cls, seq = model.encode(cu(ids), None, want_sequence=True)
    assert max_rel(cls.cpu().numpy(), out["cls"]) < 2e-5 and max_rel(seq.cpu().numpy(), out["seq"]) < 2e-5
    model.eval()
    cls2, _ = model.encode(cu(ids), cu(mask))
    assert torch.equal(cls2.detach(), cls)  # deterministic forward


def test_shared_tower_accumulates_both_passes():
    """args.share_weight: one module encodes queries and passages, gradients of both passes add"""
    cfg = bn.bert_config(num_layers=1, hidden=64, heads=1, intermediate=128, vocab=200, max_pos=64)
    params = bn.synth_params(cfg, seed=5)
    q_ids, q_mask = bn.synth_tokens(2, 8, cfg["vocab"], seed=1)
    p_ids, p_mask = bn.synth_tokens(4, 24, cfg["vocab"], seed=2, ragged=True)
    model = BiBertEncoder(None, question_model=build(cfg, params, X3))
    q, a = model(cu(q_ids), cu(q_mask), cu(p_ids), cu(p_mask))
    rng = np.random.RandomState(0)
    dq, da = rng.standard_normal((2, 64)), rng.standard_normal((4, 64))
    ((q * cu(dq, torch.float32)).sum() + (a * cu(da, torch.float32)).sum()).backward()
    ora = bn.EncoderOracle(cfg, params)
    total = None
    for ids, m, d in ((q_ids, q_mask, dq), (p_ids, p_mask, da)):
        ora.forward(ids, m)
        dc = np.zeros((ids.shape[0], 2, 64)); dc[:, -1] = d
        gr = ora.backward(d_cls=dc)
        total = gr if total is None else {k: total[k] + gr[k] for k in gr}
    assert_grads(model.question_model, total, 2e-4)
    # gradient accumulation adds into existing .grad (the reference never uses no_sync)
    before = model.question_model.embeddings.LayerNorm.weight.grad.clone()
    q, a = model(cu(q_ids), cu(q_mask), cu(p_ids), cu(p_mask))
    ((q * cu(dq, torch.float32)).sum() + (a * cu(da, torch.float32)).sum()).backward()
    assert torch.allclose(model.question_model.embeddings.LayerNorm.weight.grad, 2 * before, rtol=1e-4, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------
# fused score / loss / gradient kernel
# ---------------------------------------------------------------------------------------------------------
def test_losses_vs_reference_golden(golden_dir):
    g = np.load(golden_dir / "losses.npz")
    group = int(g["group"][0])
    t = cu(g["teacher"])

    def fresh():
        return cu(g["Q"]).requires_grad_(True), cu(g["C"]).requires_grad_(True)

    def check(q, c, loss, key, ref_loss):
        assert abs(loss.item() - ref_loss) < 1e-4, key
        loss.backward()
        assert np.abs(q.grad.cpu().numpy() - g[key + "_dQ"]).max() < 1e-5, key
        assert np.abs(c.grad.cpu().numpy() - g[key + "_dC"]).max() < 1e-5, key

    q, c = fresh()
    loss, correct = BiEncoderNllLoss().calc(q, c, [int(x) for x in g["nll_pos"]])
    assert int(correct) == int(g["nll_correct"][0])
    check(q, c, loss, "nll", g["nll_loss"][0])
    for T in (1, 2):
        q, c = fresh()
        check(q, c, simans_distill_loss(q, c, t, temperature_distill=float(T)), f"kl_T{T}", g[f"kl_T{T}_loss"][0])

    class A:
        KD_type, TEMPERATURE, CE_WEIGHT, KD_WEIGHT, LwF_WEIGHT = "KD_softmax", 4.0, 0.1, 0.9, 1.0

    q, c = fresh()
    loss, correct = CrossBERTKDLoss().calc(A, q, c, t)
    assert int(correct) == int(g["prod_correct"][0])
    check(q, c, loss, "prod", g["prod_loss"][0])
    A.KD_type, A.TEMPERATURE = "prob_loss", 1.0
    q, c = fresh()
    loss, _ = CrossBERTKDLoss().calc(A, q, c, t)
    check(q, c, loss, "prob", g["prob_loss"][0])
    # SimANS wiki retriever loss (CE-to-teacher + adversarial reward term)
    from simxns_b200.losses import simans_wiki_loss
    for lam, tag in ((0.0, "wiki_l0"), (0.3, "wiki_l3")):
        q, c = fresh()
        check(q, c, simans_wiki_loss(q, c, t, temperature_normal=2.0, adv_lambda=lam), tag, g[tag + "_loss"][0])
    # PROD BiEncoderKDLoss: DE -> DE distillation on the global matrix
    from simxns_b200.losses import BiEncoderKDLoss
    A.KD_type, A.TEMPERATURE = "KD_softmax", 4.0
    q, c = fresh()
    loss, correct = BiEncoderKDLoss().calc(A, q, c, cu(g["Qt"]), cu(g["Ct"]), [int(x) for x in g["nll_pos"]])
    assert int(correct) == int(g["bikd_correct"][0])
    check(q, c, loss, "bikd", g["bikd_loss"][0])


@pytest.mark.parametrize("GQ,group,H,q_begin,q_count", [(8, 4, 64, 0, 8), (6, 3, 128, 0, 6), (16, 8, 768, 4, 4),
                                                        (1024, 16, 768, 256, 128)])
def test_score_loss_local_slices_vs_oracle(GQ, group, H, q_begin, q_count):
    """the global matrix with this rank owning rows [q_begin, +q_count) — multi-GPU bookkeeping without the GPUs;
    the last case is BASELINE configs[2]'s full size: [1024, 16384] scores, rank 2's slice"""
    rng = np.random.RandomState(GQ + group)
    Q = (rng.standard_normal((GQ, H)) * (4.0 / np.sqrt(H))).astype(np.float32)
    Cm = (rng.standard_normal((GQ * group, H)) * (4.0 / np.sqrt(H))).astype(np.float32)
    t = (3 * rng.standard_normal((q_count, group))).astype(np.float32)
    terms = ["nll_global", "kl_local"]
    w = dict(nll_global=0.2, kl_local=1.0)
    ref, rdQ, rdC = ln.score_loss(Q, Cm, terms=terms, q_begin=q_begin, q_count=q_count, c_begin=q_begin * group,
                                  group=group, pos_idx=np.arange(GQ) * group, teacher=t, w=w)
    lib = _lib.lib()
    cfg = _lib.LossCfg(terms=_lib.LOSS_NLL_GLOBAL | _lib.LOSS_KL_LOCAL, GQ=GQ, GC=GQ * group, H=H, q_begin=q_begin,
                       q_count=q_count, c_begin=q_begin * group, c_count=q_count * group, group=group, w_nll_global=0.2,
                       w_kl_local=1.0, w_ce_local=0, w_kd_local=0, w_nll_local=0, temperature=1.0, eps=1e-7,
                       score_scale=1.0, grad_scale=1.0)
    tq, tc, tt = cu(Q), cu(Cm), cu(t)
    pos = torch.arange(0, GQ * group, group, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.simx_score_loss_workspace_bytes(C.byref(cfg)), dtype=torch.uint8, device="cuda")
    out = torch.empty(8, device="cuda")
    dQ, dC = torch.empty(q_count, H, device="cuda"), torch.empty(q_count * group, H, device="cuda")
    _lib.check(lib.simx_score_loss(C.byref(cfg), tq.data_ptr(), tc.data_ptr(), pos.data_ptr(), tt.data_ptr(),
                                   out.data_ptr(), dQ.data_ptr(), dC.data_ptr(), ws.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), "simx_score_loss")
    assert abs(out[0].item() - ref["total"]) < 1e-4
    assert abs(out[1].item() - ref["nll_global"]) < 1e-4 and abs(out[2].item() - ref["kl_local"]) < 1e-4
    assert out[6].item() == ref["n_correct"]
    assert max_rel(dQ.cpu().numpy(), rdQ) < 1e-4 and max_rel(dC.cpu().numpy(), rdC) < 1e-4


def test_score_loss_properties_full_size():
    """size-independent properties at BASELINE's full per-rank size: gradient rows of a softmax sum to zero against
    the all-ones direction; equal embeddings give log(GC); scaling grad via the loss cotangent is linear"""
    GQ, group, H = 1024, 16, 768
    q = torch.ones(GQ, H, device="cuda") * 0.03
    c = torch.ones(GQ * group, H, device="cuda") * 0.03
    loss, _ = fused_score_loss(q, c, ["nll_global"], group_size=group, gather=False)
    assert abs(loss.item() - np.log(GQ * group)) < 1e-4
    gen = torch.Generator(device="cuda").manual_seed(0)
    q = (torch.randn(GQ, H, device="cuda", generator=gen) * 0.1).requires_grad_(True)
    c = (torch.randn(GQ * group, H, device="cuda", generator=gen) * 0.1).requires_grad_(True)
    loss, _ = fused_score_loss(q, c, ["nll_global"], group_size=group, gather=False)
    (3.0 * loss).backward()
    g3 = q.grad.clone()
    q.grad = None
    loss, _ = fused_score_loss(q, c, ["nll_global"], group_size=group, gather=False)
    loss.backward()
    assert torch.allclose(g3, 3.0 * q.grad, rtol=1e-4, atol=1e-9)
    # d(loss)/dC summed over all passages equals -(1/GQ) sum_q (onehot - p) q = 0 contribution identity:
    # sum_c dC[c] == sum_q (sum_c dS[q,c]) Q[q] and each dS row sums to zero
    assert c.grad.sum(0).abs().max().item() < 1e-5


def test_bad_arguments_raise_with_message():
    q = torch.randn(4, 64, device="cuda")
    c = torch.randn(9, 64, device="cuda")
    with pytest.raises(RuntimeError, match="c_count == q_count"):
        fused_score_loss(q, c, ["kl_local"], teacher=torch.randn(4, 2, device="cuda"), group_size=2, gather=False)
    cfg = bn.bert_config(num_layers=1, hidden=64, heads=1, intermediate=64, vocab=50, max_pos=16)
    m = build(cfg, bn.synth_params(cfg, seed=0), BF16)
    with pytest.raises(ValueError, match="exceeds"):
        m.encode(torch.ones(1, 17, dtype=torch.long, device="cuda"))


# ---------------------------------------------------------------------------------------------------------
# LEAD mutual distillation (simx_pair_distill) vs the real LEAD distill_loss
# ---------------------------------------------------------------------------------------------------------
def test_lead_distill_vs_reference_golden(golden_dir):
    from simxns_b200.losses import lead_distill_loss

    g = np.load(golden_dir / "lead_distill.npz")
    sel_db, sel_de = [int(x) for x in g["sel_db"]], [int(x) for x in g["sel_de"]]
    for tag in ("both", "frozen"):
        train_de, reweight, T, Tl, p_db, p_de, p_dis, p_layer = [float(x) for x in g[tag + "_cfg"]]
        t = {n: cu(g[n]).requires_grad_(True) for n in ("q_db", "c_db", "q_de", "c_de", "qL_db", "cL_db", "qL_de", "cL_de")}
        loss, parts = lead_distill_loss(t["q_db"], t["c_db"], t["q_de"], t["c_de"], t["qL_db"][:, :, 0], t["cL_db"][:, :, 0],
                                        t["qL_de"][:, :, 0], t["cL_de"][:, :, 0], selected_db=sel_db, selected_de=sel_de,
                                        temperature=T, layer_temperature=Tl, para_db=p_db, para_de=p_de, para_dis=p_dis,
                                        para_layer=p_layer, reweight=bool(reweight), train_de=bool(train_de), gather=False)
        assert abs(loss.item() - g[tag + "_loss"][0]) < 1e-4
        loss.backward()
        for n in t:
            ref = g[f"{tag}_d_{n}"]
            got = t[n].grad.cpu().numpy() if t[n].grad is not None else np.zeros_like(ref)
            assert np.abs(got - ref).max() < 1e-3 * max(1e-3, np.abs(ref).max()), (tag, n)


def test_pair_distill_local_slices_vs_oracle():
    rng = np.random.RandomState(4)
    GQ, group, H = 64, 8, 256
    mk = lambda n: (rng.standard_normal((n, H)) * (3.0 / np.sqrt(H))).astype(np.float32)
    Qa, Ca, Qb, Cb = mk(GQ), mk(GQ * group), mk(GQ), mk(GQ * group)
    kw = dict(w_nll_a=1.0, w_nll_b=0.5, w_kl_b2a=1.5, w_kl_a2b=0.25, temperature=2.0)
    q0, qn = 16, 16
    ref, *rg = ln.pair_distill(Qa, Ca, Qb, Cb, pos_idx=np.arange(GQ) * group, q_begin=q0, q_count=qn, c_begin=q0 * group,
                               c_count=qn * group, **kw)
    lib = _lib.lib()
    cfg = _lib.PairCfg(GQ=GQ, GC=GQ * group, H=H, q_begin=q0, q_count=qn, c_begin=q0 * group, c_count=qn * group,
                       grad_scale=1.0, **kw)
    ts = [cu(x) for x in (Qa, Ca, Qb, Cb)]
    pos = torch.arange(0, GQ * group, group, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.simx_pair_distill_workspace_bytes(GQ, GQ * group, H), dtype=torch.uint8, device="cuda")
    out = torch.empty(8, device="cuda")
    gs = [torch.empty(qn, H, device="cuda"), torch.empty(qn * group, H, device="cuda"), torch.empty(qn, H, device="cuda"),
          torch.empty(qn * group, H, device="cuda")]
    _lib.check(lib.simx_pair_distill(C.byref(cfg), *[t.data_ptr() for t in ts], pos.data_ptr(), out.data_ptr(),
                                     *[t.data_ptr() for t in gs], ws.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "simx_pair_distill")
    assert abs(out[0].item() - ref["total"]) < 1e-4
    for got, want in zip(gs, rg):
        assert max_rel(got.cpu().numpy(), want) < 1e-4


def test_lead_end_to_end_step_vs_oracle():
    """LEAD distillation step: 4-layer BERT teacher -> 2-layer DistilBERT student, both shared-weight dual
    encoders, response + per-layer score distillation with lambda re-weighting; gradients enter the towers at
    intermediate layers' [CLS] rows.  Checked end to end against the numpy oracle."""
    from simxns_b200 import lead_models as lm
    from simxns_b200.losses import lead_distill_loss

    cfg_t = bn.bert_config(num_layers=4, hidden=64, heads=1, intermediate=128, vocab=200, max_pos=64)
    cfg_s = bn.bert_config(num_layers=2, hidden=64, heads=1, intermediate=128, vocab=200, max_pos=64, arch="distilbert")
    p_t, p_s = bn.synth_params(cfg_t, seed=51), bn.synth_params(cfg_s, seed=52)
    # keep the problem well conditioned: with unit LayerNorm gains |q.c| reaches ~65 on random weights, every softmax
    # row is one-hot to 1e-20 and the gradients are pure cancellation (fp32 round-off of the scores dominates them)
    for p in (p_t, p_s):
        for k in p:
            if k.endswith("LayerNorm.weight") or k.endswith("layer_norm.weight"):
                p[k] = (p[k] * 0.3).astype(np.float32)
    nq, group = 4, 2
    q_ids, q_mask = bn.synth_tokens(nq, 12, 200, seed=53, ragged=True)
    d_ids, d_mask = bn.synth_tokens(nq * group, 40, 200, seed=54, ragged=True, min_len=6)
    sel_s, sel_t = [0, 1], [1, 3]
    hp = dict(temperature=2.0, layer_temperature=0.5, para_db=1.0, para_de=0.7, para_dis=1.5)

    # ---- oracle
    ot, os_ = bn.EncoderOracle(cfg_t, p_t), bn.EncoderOracle(cfg_s, p_s)
    outs = {}
    for name, ora, ids, m in (("tq", ot, q_ids, q_mask), ("td", ot, d_ids, d_mask), ("sq", os_, q_ids, q_mask),
                              ("sd", os_, d_ids, d_mask)):
        outs[name] = ora.forward(ids, m)["cls"]
    total, gr = ln.lead_de_db_loss(outs["sq"][:, -1], outs["sd"][:, -1], outs["tq"][:, -1], outs["td"][:, -1],
                                   [outs["sq"][:, l + 1] for l in sel_s], [outs["sd"][:, l + 1] for l in sel_s],
                                   [outs["tq"][:, l + 1] for l in sel_t], [outs["td"][:, l + 1] for l in sel_t],
                                   group=group, reweight=True, train_de=True, **hp)

    def tower_grads(cfg, params, passes):
        tot = None
        for ids, m, d_final, layer_sel, d_layers in passes:
            ora = bn.EncoderOracle(cfg, params)
            ora.forward(ids, m)
            d_cls = np.zeros((ids.shape[0], cfg["num_layers"] + 1, cfg["hidden"]))
            d_cls[:, -1] += d_final
            for l, d in zip(layer_sel, d_layers):
                d_cls[:, l + 1] += d
            g = ora.backward(d_cls=d_cls)
            tot = g if tot is None else {k: tot[k] + g[k] for k in g}
        return tot

    ref_s = tower_grads(cfg_s, p_s, [(q_ids, q_mask, gr["q_db"], sel_s, gr["qL_db"]), (d_ids, d_mask, gr["c_db"], sel_s, gr["cL_db"])])
    ref_t = tower_grads(cfg_t, p_t, [(q_ids, q_mask, gr["q_de"], sel_t, gr["qL_de"]), (d_ids, d_mask, gr["c_de"], sel_t, gr["cL_de"])])

    # ---- B200
    student = lm.BiBertEncoder(None, question_model=lm.HFDistilBertEncoder(ec_from(cfg_s), precision=X3))
    teacher = lm.BiBertEncoder(None, question_model=lm.HFBertEncoder(ec_from(cfg_t), precision=X3))
    student.question_model.load_state_dict({k: torch.from_numpy(v) for k, v in p_s.items()})
    teacher.question_model.load_state_dict({k: torch.from_numpy(v) for k, v in p_t.items()})
    student.cuda(), teacher.cuda()
    args = (cu(q_ids), cu(q_mask), cu(d_ids), cu(d_mask))
    q_db, d_db, qa_db, da_db, _ = student(*args)
    q_de, d_de, qa_de, da_de, _ = teacher(*args)
    assert qa_db.shape == (nq, 2, 1, 64) and da_de.shape == (nq * group, 4, 1, 64)
    loss, parts = lead_distill_loss(q_db, d_db, q_de, d_de, qa_db[:, :, 0], da_db[:, :, 0], qa_de[:, :, 0], da_de[:, :, 0],
                                    selected_db=sel_s, selected_de=sel_t, reweight=True, train_de=True, gather=False, **hp)
    assert abs(loss.item() - total) < 1e-4
    loss.backward()
    assert_grads(student.question_model, ref_s, 1e-3)
    assert_grads(teacher.question_model, ref_t, 1e-3)


def test_lead_ce_teacher_distill_vs_reference_golden(golden_dir):
    from simxns_b200.losses import lead_ce_distill_loss

    g = np.load(golden_dir / "lead_distill.npz")
    sel_db, sel_t = [int(x) for x in g["sel_db"]], [int(x) for x in g["sel_de"]]
    T, Tl, p_db, p_ce, p_dis = [float(x) for x in g["ce_cfg"]]
    t = {n: cu(g[n]).requires_grad_(True) for n in ("q_db", "c_db", "qL_db", "cL_db", "ce_logits", "ce_layers")}
    loss, parts = lead_ce_distill_loss(t["q_db"], t["c_db"], t["ce_logits"], t["qL_db"][:, :, 0], t["cL_db"][:, :, 0],
                                       t["ce_layers"], selected_db=sel_db, selected_ce=sel_t, temperature=T,
                                       layer_temperature=Tl, para_db=p_db, para_ce=p_ce, para_dis=p_dis, train_ce=True)
    assert abs(loss.item() - g["ce_loss"][0]) < 1e-4
    loss.backward()
    for n in t:
        ref = g[f"ce_d_{n}"]
        assert np.abs(t[n].grad.cpu().numpy() - ref).max() < 1e-3 * max(1e-3, np.abs(ref).max()), n


# ---------------------------------------------------------------------------------------------------------
# checkpoint layout + training-loop contract
# ---------------------------------------------------------------------------------------------------------
def test_checkpoint_round_trip_with_reference_model(tmp_path):
    """north_star "same checkpoint layout": a CheckpointState file written from this model loads into the reference's
    HF-based BiBertEncoder (oracle/ref_torch.py port of SimANS/model/models.py:58-118) and vice versa, and both produce the
    same embeddings (SimANS/utils/dpr_utils.py:22-24,73-77; co_training_marco_train.py:310-326)."""
    import collections

    from oracle import ref_torch as rt

    CheckpointState = collections.namedtuple("CheckpointState", ['model_dict', 'optimizer_dict', 'scheduler_dict', 'offset',
                                                                 'epoch', 'encoder_params'])
    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=64)
    ours = BiBertEncoder(None, question_model=build(cfg, bn.synth_params(cfg, seed=61), X3),
                         ctx_model=build(cfg, bn.synth_params(cfg, seed=62), X3))
    opt = torch.optim.AdamW(ours.parameters(), lr=1e-3)
    state = CheckpointState(ours.state_dict(), opt.state_dict(), {}, 7, 0, {})
    torch.save(state._asdict(), tmp_path / "checkpoint-7")
    saved = torch.load(tmp_path / "checkpoint-7", map_location="cpu", weights_only=False)
    assert set(saved) == {'model_dict', 'optimizer_dict', 'scheduler_dict', 'offset', 'epoch', 'encoder_params'}
    ref = rt.BiBertEncoderRef(cfg)
    missing, unexpected = ref.load_state_dict(saved["model_dict"], strict=False)
    assert not unexpected and all("position_ids" in m or "token_type_ids" in m for m in missing)
    ref.eval()
    q_ids, q_mask = bn.synth_tokens(3, 16, cfg["vocab"], seed=1, ragged=True)
    p_ids, p_mask = bn.synth_tokens(6, 40, cfg["vocab"], seed=2, ragged=True)
    with torch.no_grad():
        rq, rp = ref(torch.from_numpy(q_ids), torch.from_numpy(q_mask), torch.from_numpy(p_ids), torch.from_numpy(p_mask))
        oq, op = ours(cu(q_ids), cu(q_mask), cu(p_ids), cu(p_mask))
    assert (oq.cpu() - rq).abs().max() < 1e-4 and (op.cpu() - rp).abs().max() < 1e-4
    # the other direction: a reference checkpoint (with the legacy position_ids buffer) into a fresh B200 model
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    sd["question_model.embeddings.position_ids"] = torch.arange(64)[None]
    fresh = BiBertEncoder(None, question_model=HFBertEncoder(ec_from(cfg), precision=X3),
                          ctx_model=HFBertEncoder(ec_from(cfg), precision=X3))
    fresh.load_state_dict(sd, strict=False)
    fresh.cuda()
    with torch.no_grad():
        fq, _ = fresh(cu(q_ids), cu(q_mask), cu(p_ids), cu(p_mask))
    assert (fq.cpu() - rq).abs().max() < 1e-4


def test_training_loop_contract_adamw_clip_accumulate():
    """the retriever loop of co_training_marco_train.py:186-263 on this model: AdamW with the bias/LayerNorm no-decay
    grouping, gradient accumulation, clip_grad_norm_, optimizer.step, zero_grad; weights re-staged after every step
    (the loss must go down on a fixed batch)."""
    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=64)
    model = BiBertEncoder(None, question_model=build(cfg, bn.synth_params(cfg, seed=71), BF16),
                          ctx_model=build(cfg, bn.synth_params(cfg, seed=72), BF16))
    no_decay = ['bias', 'LayerNorm.weight']
    groups = [{'params': [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], 'weight_decay': 0.01},
              {'params': [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], 'weight_decay': 0.0}]
    opt = torch.optim.AdamW(groups, lr=2e-4, eps=1e-8)
    nq, group = 8, 4
    q_ids, q_mask = bn.synth_tokens(nq, 16, cfg["vocab"], seed=3, ragged=True)
    p_ids, p_mask = bn.synth_tokens(nq * group, 32, cfg["vocab"], seed=4, ragged=True)
    teacher = cu((3 * np.random.RandomState(5).standard_normal((nq, group))).astype(np.float32))
    batch = (cu(q_ids), cu(q_mask), cu(p_ids), cu(p_mask))
    losses, accum = [], 2
    for step in range(12):
        q, a = model(*batch)
        loss = simans_distill_loss(q, a, teacher, temperature_distill=1.0) / accum
        loss.backward()
        if (step + 1) % accum == 0:
            gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 2.0)
            assert torch.isfinite(gn)
            opt.step()
            model.zero_grad()
            losses.append(loss.item() * accum)
    assert losses[-1] < losses[0] - 0.05, losses


# ---------------------------------------------------------------------------------------------------------
# the encoder at the full per-GPU size of BASELINE configs[2] (bert-base, 2048 passages x 128 tokens, bf16):
# size-independent properties — the oracle cannot run this size
# ---------------------------------------------------------------------------------------------------------
def test_encoder_full_size_properties_bert_base():
    """(1) batch-permutation equivariance, bit-exact: a sequence's embedding does not depend on where it sits in the
    batch (tile scheduling, CTA pairs, persistent loops); (2) padding invariance: appending masked pad tokens to a
    sequence leaves its [CLS] embedding unchanged up to bf16 round-off of the different tile shapes; (3) the gradient is
    linear in the cotangent and the gradient of padded positions' word rows is exactly zero."""
    ec = EncoderConfig()  # bert-base
    torch.manual_seed(0)
    enc = HFBertEncoder(ec, precision=BF16).cuda()
    gen = torch.Generator(device="cuda").manual_seed(1)
    B, S = 2048, 128
    ids = torch.randint(1000, ec.vocab, (B, S), device="cuda", generator=gen)
    ids[:, 0] = 101
    lens = torch.randint(16, S + 1, (B,), device="cuda", generator=gen)
    mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).long()
    ids = ids * mask
    with torch.no_grad():
        cls, _ = enc.encode(ids, mask)
        perm = torch.randperm(B, device="cuda", generator=gen)
        cls_p, _ = enc.encode(ids[perm], mask[perm])
    assert torch.equal(cls_p, cls[perm])
    assert torch.isfinite(cls).all()
    # padding invariance on a slice (S 96 -> 128 with masked tail)
    short = lens <= 96
    sub = short.nonzero().flatten()[:256]
    with torch.no_grad():
        cls_s, _ = enc.encode(ids[sub, :96].contiguous(), mask[sub, :96].contiguous())
    rel = ((cls_s[:, -1] - cls[sub, -1]).abs().max() / cls[sub, -1].abs().max()).item()
    assert rel < 3e-2, rel
    # backward on a 256-sequence slice: linearity in the cotangent, zero gradient for the pad row (id 0 is only ever masked)
    sl = slice(0, 256)
    d = torch.randn(256, ec.hidden, device="cuda", generator=gen)
    out, _ = enc.encode(ids[sl], mask[sl])
    (out[:, -1] * d).sum().backward()
    g1 = enc.encoder.layer[5].intermediate.dense.weight.grad.clone()
    gw = enc.embeddings.word_embeddings.weight.grad.clone()
    enc.zero_grad()
    out, _ = enc.encode(ids[sl], mask[sl])
    (out[:, -1] * (2.0 * d)).sum().backward()
    g2 = enc.encoder.layer[5].intermediate.dense.weight.grad
    assert ((g2 - 2.0 * g1).abs().max() / g1.abs().max()).item() < 2e-2  # bf16 activations: rounding is not linear
    used = torch.zeros(ec.vocab, dtype=torch.bool, device="cuda")
    used[ids[sl].flatten()] = True
    assert float(gw[~used].abs().max()) == 0.0  # rows of tokens that never occur get exactly nothing


def test_bf16_mode_tracks_fp32_class_mode_at_bert_base_width():
    """the throughput mode against the fp32-class mode of the SAME kernels (itself pinned to the reference fixtures and
    the oracle above) at the real width — H 768, 12 heads, I 3072, 512 passages x 128 tokens, 2 layers — where the oracle
    is too slow: [CLS] embeddings within 3e-2 of their scale, every parameter gradient within 8e-2 of its tensor's scale
    (the tolerances of the small-shape bf16 tests; SURVEY D6: the reference itself moves 3 % under bf16 autocast)"""
    ec = EncoderConfig(num_layers=2)
    torch.manual_seed(0)
    ref = HFBertEncoder(ec, precision=X3).cuda()
    fast = HFBertEncoder(ec, precision=BF16).cuda()
    fast.load_state_dict(ref.state_dict())
    gen = torch.Generator(device="cuda").manual_seed(5)
    B, S = 512, 128
    ids = torch.randint(1000, ec.vocab, (B, S), device="cuda", generator=gen)
    ids[:, 0] = 101
    lens = torch.randint(8, S + 1, (B,), device="cuda", generator=gen)
    mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).long()
    ids = ids * mask
    d = torch.randn(B, ec.hidden, device="cuda", generator=gen)
    outs = []
    for m in (ref, fast):
        cls, _ = m.encode(ids, mask)
        (cls[:, -1] * d).sum().backward()
        outs.append(cls[:, -1].detach())
    assert ((outs[1] - outs[0]).abs().max() / outs[0].abs().max()).item() < 3e-2
    scales = [p.grad.abs().max().item() for p in ref.parameters()]
    floor = 1e-3 * float(np.median([s for s in scales if s > 0]))
    worst = ("", 0.0)
    for (n, p), (_, q) in zip(ref.named_parameters(), fast.named_parameters()):
        if "key.bias" in n:  # true-zero gradient (softmax shift invariance): round-off only in both modes
            continue
        e = ((q.grad - p.grad).abs().max() / max(p.grad.abs().max().item(), floor)).item()
        if e > worst[1]:
            worst = (n, e)
    assert worst[1] < 8e-2, worst


def test_bf16_mode_deviates_no_more_than_the_reference_under_bf16_autocast(golden_dir):
    """The throughput mode pinned to the reference's OWN reduced-precision behaviour: tests/golden/biencoder_cfg1_bf16_autocast.npz
    holds the cfg1 step (BASELINE configs[0]/[1]: bert-base, 2 queries, 1+1 passages, SimANS KL loss) run by the reference's
    classes under torch.autocast(bfloat16) (oracle/make_golden.py cfg1_bf16).  For every quantity — embeddings, loss,
    per-tensor gradient norms of all 398 tensors, four complete gradient tensors — the distance of the bf16 kernels from the
    reference's fp32 result is compared with the distance of the reference's own bf16 run from it.

    Same class, not identical: autocast rounds only the matmul operands / outputs and keeps the residual stream and the
    LayerNorm outputs in fp32, while these kernels STORE every activation in bf16 (that is what halves the HBM traffic), so the
    forward deviation is a small multiple of the reference's (measured 2.6x: 1.5 % vs 0.6 % of the embedding scale), and this
    loss (KL at 0.03, p_s - p_t nearly cancelling) amplifies it into the gradients (median per-tensor norm error 9 % vs 4.5 %;
    the worst tensor 12 % vs the reference's 22 %; worst ratio of a complete tensor 3.17x, identical on every run).  Bound: 3.5x the reference's own deviation (+ a small floor); the numbers
    are printed and recorded in DESIGN.md.  Training runs that need the reference's fp32 behaviour use SIMX_PREC_BF16X3."""
    g32 = np.load(golden_dir / "biencoder_cfg1.npz")
    g16 = np.load(golden_dir / "biencoder_cfg1_bf16_autocast.npz")
    cfg = bn.bert_config()
    pq, pc = bn.synth_params(cfg, seed=0, perturb=False), bn.synth_params(cfg, seed=1, perturb=False)
    model = BiBertEncoder(None, question_model=build(cfg, pq, BF16), ctx_model=build(cfg, pc, BF16))
    q, a = model(cu(g32["q_ids"]), cu(g32["q_mask"]), cu(g32["p_ids"]), cu(g32["p_mask"]))
    loss = simans_distill_loss(q, a, cu(g32["teacher"]), temperature_distill=1.0)
    loss.backward()
    ours = dict(q_embs=q.detach().cpu().numpy(), a_embs=a.detach().cpu().numpy())
    report, checks = {}, []
    for k in ("q_embs", "a_embs"):
        d_ref = float(np.abs(g16[k] - g32[k]).max() / np.abs(g32[k]).max())
        d_us = float(np.abs(ours[k] - g32[k]).max() / np.abs(g32[k]).max())
        report[k] = (d_us, d_ref)
        checks.append((k, d_us, d_ref, 1e-3))
    d_ref, d_us = abs(float(g16["loss"][0] - g32["loss"][0])), abs(loss.item() - float(g32["loss"][0]))
    report["loss"] = (d_us, d_ref)
    checks.append(("loss", d_us, d_ref, 1e-3))
    grads = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters()}
    for k in [k for k in g32.files if k.startswith("grad/")]:  # complete tensors
        n = k[5:]
        scale = np.abs(g32[k]).max()
        d_ref, d_us = float(np.abs(g16[k] - g32[k]).max() / scale), float(np.abs(grads[n] - g32[k]).max() / scale)
        report[n] = (d_us, d_ref)
        checks.append((n, d_us, d_ref, 5e-3))
    # gradient norms of every tensor (true-zero gradients — key biases, pooler — excluded: both runs hold round-off only)
    norms = [(k[6:], g32[k][0], g16[k][0]) for k in g32.files if k.startswith("gstat/")]
    big = max(v for _, v, _ in norms)
    rel_us, rel_ref = [], []
    for n, v32, v16 in norms:
        if v32 < 1e-6 * big:
            continue
        rel_ref.append(abs(v16 - v32) / v32)
        rel_us.append(abs(float(np.linalg.norm(grads[n].astype(np.float64))) - v32) / v32)
    rel_us, rel_ref = np.asarray(rel_us), np.asarray(rel_ref)
    report["grad-norm median"] = (float(np.median(rel_us)), float(np.median(rel_ref)))
    report["grad-norm max"] = (float(rel_us.max()), float(rel_ref.max()))
    checks.append(("grad-norm median", float(np.median(rel_us)), float(np.median(rel_ref)), 1e-3))
    checks.append(("grad-norm max", float(rel_us.max()), float(rel_ref.max()), 1e-2))
    print("deviation from the reference's fp32 result (bf16 kernels, reference under bf16 autocast):")
    for k, v in report.items():
        print(f"  {k:70s} {v[0]:.4e}  {v[1]:.4e}  x{v[0] / max(v[1], 1e-12):.2f}")
    bad = [(k, u, r) for k, u, r, floor in checks if u > 3.5 * r + floor]
    assert not bad, bad
