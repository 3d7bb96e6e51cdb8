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
    return cfg, params, ids, mask, head_w, head_b, out, loss, grads


@pytest.mark.parametrize("precision,tol_h,tol_l,tol_g", [(X3, 2e-5, 1e-5, 2e-4), (BF16, 5e-2, 5e-2, 1e-1)])
def test_ernie_large_width_teacher_step_vs_oracle(large_case, precision, tol_h, tol_l, tol_g):
    cfg, params, ids, mask, head_w, head_b, out, loss_ref, grads = large_case
    enc = build(cfg, params, precision)
    model = Reranker(enc, 1024).cuda()
    with torch.no_grad():
        model.qa_classifier.weight.copy_(cu(head_w))
        model.qa_classifier.bias.copy_(cu(head_b))
    logits = model(cu(ids).view(2, 3, 160), cu(mask).view(2, 3, 160))
    cls = enc.last_cls_layers.detach().cpu().numpy()
    assert max_rel(cls, out["cls"]) < tol_h
    loss = torch.nn.functional.cross_entropy(logits, torch.zeros(2, dtype=torch.long, device="cuda"))
    assert abs(loss.item() - loss_ref) < tol_l * max(1.0, abs(loss_ref))
    loss.backward()
    # true-zero gradients: the head bias and the last LayerNorm's bias both receive sum_b (softmax - onehot)_b = 0; the last
    # FFN-output bias is the same sum pushed through six different LayerNorm Jacobians — 1e-3 of the largest gradient, all
    # cancellation — so the bf16 mode is judged on the global scale there (the fp32-class mode holds it to 2e-4 of itself)
    zero_true = ("qa_classifier.bias", "encoder.encoder.layer.1.output.LayerNorm.bias")
    if precision == X3:
        assert_grads(model, grads, tol_g, zero_true=zero_true)
        return
    # bf16 mode: this loss (softmax - onehot over 3 candidates, 2 queries) makes most bias / LayerNorm gradients sums of
    # cancelling terms, 1e-2 .. 1e-3 of the largest gradient; bf16 activations resolve those to ~30 % of themselves (the
    # fp32-class mode above holds them to 2e-4).  Judge every tensor on max(own scale, 5 % of the largest gradient).
    gmax = max(np.abs(v).max() for v in grads.values())
    worst = ("", 0.0)
    for n, p in model.named_parameters():
        r = grads[n]
        # the token-type row and the embedding LayerNorm affine sum the (bf16-rounded) cotangent over all 960 tokens
        floor = (0.25 if "embeddings.token_type" in n or "embeddings.LayerNorm" in n else 5e-2) * gmax
        e = float(np.abs(p.grad.cpu().numpy() - r).max() / max(np.abs(r).max(), floor))
        if e > worst[1]:
            worst = (n, e)
    assert worst[1] < tol_g, worst


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
