"""B200 (`-m gpu`): the teacher shape of the shipped SimANS recipes — `--teacher_model_type=nghuyong/ernie-2.0-large-en`
(SimANS/train_MS_Pas_AR2.sh:13; Reranker over HFBertEncoder, SimANS/model/models.py:638-659): hidden 1024, 16 heads,
intermediate 4096 — at 2 layers and the cross-encoder length 160, forward + backward against the numpy oracle in both
precision modes (round 1 only allowed this width, nothing ran at it)."""
import numpy as np
import pytest
import torch

from oracle import bert_numpy as bn
from simxns_b200 import _lib
from simxns_b200.models import Reranker
from test_gpu_parity import X3, BF16, assert_grads, build, cu, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    _lib.require_device()


@pytest.fixture(scope="module")
def large_case():
    cfg = bn.bert_config(num_layers=2, hidden=1024, heads=16, intermediate=4096, vocab=2000, max_pos=512)
    params = bn.synth_params(cfg, seed=1024)
    ids, mask = bn.synth_tokens(6, 160, cfg["vocab"], seed=160, ragged=True, min_len=20)
    ora = bn.EncoderOracle(cfg, params)
    out = ora.forward(ids, mask)
    rng = np.random.RandomState(4096)
    head_w = (0.02 * rng.standard_normal((1, 1024))).astype(np.float32)
    head_b = np.zeros(1, np.float32)
    logits = (out["cls"][:, -1] @ head_w.T.astype(np.float64) + head_b).reshape(2, 3)
    # cross-entropy with the positive in column 0 (co_training_marco_train.py:225-245)
    z = logits - logits.max(1, keepdims=True)
    p = np.exp(z) / np.exp(z).sum(1, keepdims=True)
    loss = float(-np.log(p[:, 0]).mean())
    dlog = p.copy()
    dlog[:, 0] -= 1.0
    dlog /= 2
    d_cls = np.zeros(out["cls"].shape)
    d_cls[:, -1] = dlog.reshape(6, 1) @ head_w.astype(np.float64)
    grads = ora.backward(d_cls=d_cls)
    grads = {"encoder." + k: v for k, v in grads.items()}
    grads["qa_classifier.weight"] = dlog.reshape(6, 1).T @ out["cls"][:, -1]
    grads["qa_classifier.bias"] = dlog.sum().reshape(1)
    d_rand = np.random.RandomState(7).standard_normal(out["cls"].shape)  # a cotangent on every layer's [CLS] row
    grads_rand = ora.backward(d_cls=d_rand)
    return cfg, params, ids, mask, head_w, head_b, out, loss, grads, d_rand, grads_rand


def test_ernie_large_width_teacher_step_vs_oracle(large_case):
    """Reranker CE step (co_training_marco_train.py:225-245) in the fp32-class mode at the north_star tolerances"""
    cfg, params, ids, mask, head_w, head_b, out, loss_ref, grads, _, _ = large_case
    enc = build(cfg, params, X3)
    model = Reranker(enc, 1024).cuda()
    with torch.no_grad():
        model.qa_classifier.weight.copy_(cu(head_w))
        model.qa_classifier.bias.copy_(cu(head_b))
    logits = model(cu(ids).view(2, 3, 160), cu(mask).view(2, 3, 160))
    cls = enc.last_cls_layers.detach().cpu().numpy()
    assert max_rel(cls, out["cls"]) < 2e-5
    loss = torch.nn.functional.cross_entropy(logits, torch.zeros(2, dtype=torch.long, device="cuda"))
    assert abs(loss.item() - loss_ref) < 1e-5 * max(1.0, abs(loss_ref))
    loss.backward()
    # true-zero gradients: the head bias and the last LayerNorm's bias both receive sum_b (softmax - onehot)_b = 0
    assert_grads(model, grads, 2e-4, zero_true=("qa_classifier.bias", "encoder.encoder.layer.1.output.LayerNorm.bias"))


@pytest.mark.parametrize("precision,tol_h,tol_g", [(X3, 2e-5, 2e-4), (BF16, 5e-2, 1e-1)])
def test_ernie_large_width_tower_vs_oracle_random_cotangent(large_case, precision, tol_h, tol_g):
    """both precision modes at H 1024 / I 4096 / S 160 under a generic cotangent on every layer's [CLS] row (the CE loss
    above makes most bias / LayerNorm / token-type gradients sums of cancelling terms — fine for the fp32-class mode, but no
    yardstick for bf16 activations)"""
    cfg, params, ids, mask, _, _, out, _, _, d_rand, grads_rand = large_case
    enc = build(cfg, params, precision)
    cls, _ = enc.encode(cu(ids), cu(mask))
    assert max_rel(cls.detach().cpu().numpy(), out["cls"]) < tol_h
    (cls * cu(d_rand, torch.float32)).sum().backward()
    assert_grads(enc, grads_rand, tol_g)


def test_ernie_large_width_with_checkpointing_and_dropout_runs_and_matches_itself():
    """the configuration the shipped script actually trains the teacher in: dropout 0.1 + --gradient_checkpointing"""
    cfg = bn.bert_config(num_layers=2, hidden=1024, heads=16, intermediate=4096, vocab=2000, max_pos=512)
    params = bn.synth_params(cfg, seed=5)
    ids, mask = bn.synth_tokens(4, 160, cfg["vocab"], seed=6, ragged=True, min_len=20)
    d = torch.randn(4, 1024, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    res = []
    for ck in (False, True):
        m = build(cfg, params, BF16).train()
        m.hidden_dropout_prob = m.attention_probs_dropout_prob = 0.1
        m.forced_dropout_seed = 99
        if ck:
            m.gradient_checkpointing_enable()
        cls, _ = m.encode(cu(ids), cu(mask))
        (cls[:, -1] * d).sum().backward()
        res.append((cls.detach().clone(), m.encoder.layer[0].intermediate.dense.weight.grad.clone()))
    assert torch.equal(res[0][0], res[1][0])
    e = (res[0][1] - res[1][1]).abs().max().item() / res[0][1].abs().max().item()
    assert e < 2e-2, e
