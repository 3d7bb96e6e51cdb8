"""B200 (`-m gpu`): the holes VERDICT r01 listed — MASTER's `normal + 0.2 * InfoNCE` as the exact term pair
(MASTER/finetune/MS/co_training_model.py:249-269: CE-to-teacher on the local block + 0.2 x in-batch NLL), and LEAD's
Reranker with per-layer relevance logits (LEAD/models.py:249-267), forward and backward against the oracle."""
import numpy as np
import pytest
import torch

from oracle import bert_numpy as bn
from oracle import losses_numpy as ln
from simxns_b200 import _lib, lead_models
from simxns_b200.encoder import EncoderConfig
from simxns_b200.losses import fused_score_loss
from test_gpu_parity import X3, assert_grads, cu, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    _lib.require_device()


def test_master_ce_local_plus_02_nll_global_pair():
    rng = np.random.RandomState(0)
    nq, group, H = 24, 16, 768
    Q = rng.standard_normal((nq, H)).astype(np.float32) * 0.3
    Cm = rng.standard_normal((nq * group, H)).astype(np.float32) * 0.3
    teacher = (3 * rng.standard_normal((nq, group))).astype(np.float32)
    w = dict(ce_local=1.0, nll_global=0.2)
    ref, dQ, dC = ln.score_loss(Q, Cm, terms=["ce_local", "nll_global"], group=group, pos_idx=np.arange(nq) * group,
                                teacher=teacher, w=w, temperature=1.0)
    q, c = cu(Q).requires_grad_(True), cu(Cm).requires_grad_(True)
    loss, detail = fused_score_loss(q, c, ["ce_local", "nll_global"], teacher=cu(teacher), weights=w, group_size=group,
                                    gather=False)
    assert abs(loss.item() - ref["total"]) < 1e-4
    assert abs(detail[3].item() - ref["ce_local"]) < 1e-4 and abs(detail[1].item() - ref["nll_global"]) < 1e-4
    loss.backward()
    assert np.abs(q.grad.cpu().numpy() - dQ).max() < 1e-5 and np.abs(c.grad.cpu().numpy() - dC).max() < 1e-5


def test_lead_reranker_per_layer_logits_forward_backward():
    cfg = bn.bert_config(num_layers=3, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=64)
    params = bn.synth_params(cfg, seed=17)
    N, M, S = 3, 4, 40
    ids, mask = bn.synth_tokens(N * M, S, cfg["vocab"], seed=5, ragged=True, min_len=6)
    enc = lead_models.HFBertEncoder(EncoderConfig(num_layers=3, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=64),
                                    precision=X3)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    model = lead_models.Reranker(enc, 128).cuda()
    rng = np.random.RandomState(3)
    hw = (0.05 * rng.standard_normal((1, 128))).astype(np.float32)
    hb = np.asarray([0.01], np.float32)
    with torch.no_grad():
        model.qa_classifier.weight.copy_(cu(hw))
        model.qa_classifier.bias.copy_(cu(hb))
    logits, attn, layer_logits = model(cu(ids).view(N, M, S), cu(mask).view(N, M, S))
    assert attn is None and tuple(logits.shape) == (N, M) and tuple(layer_logits.shape) == (3, N, M)

    ora = bn.EncoderOracle(cfg, params)
    out = ora.forward(ids, mask)
    cls = out["cls"]                                    # [NM, L+1, H]
    want_layers = (cls[:, 1:] @ hw.T.astype(np.float64) + hb).transpose(1, 0, 2).reshape(3, N, M)
    assert max_rel(layer_logits.detach().cpu().numpy(), want_layers) < 2e-5
    assert max_rel(logits.detach().cpu().numpy(), want_layers[-1]) < 2e-5
    assert torch.equal(logits, layer_logits[-1])

    # a loss that reads every layer's logits (what LEAD's CE-teacher distillation does): cotangents on every layer's [CLS]
    cw = rng.standard_normal((3, N, M))
    (layer_logits * cu(cw, torch.float32)).sum().backward()
    d_logit = cw.reshape(3, N * M).T                    # [NM, L]
    d_cls = np.zeros_like(cls)
    d_cls[:, 1:] = d_logit[:, :, None] * hw.astype(np.float64)[None]
    grads = {"encoder." + k: v for k, v in ora.backward(d_cls=d_cls).items()}
    grads["qa_classifier.weight"] = np.einsum("nl,nlh->h", d_logit, cls[:, 1:])[None]
    grads["qa_classifier.bias"] = d_logit.sum().reshape(1)
    assert_grads(model, grads, 2e-4)


def test_inputs_embeds_and_forward_adv_paths_vs_oracle():
    """`inputs_embeds=` through the kernels (simx_encoder_forward_embeds / _backward_embeds) and the three forward_adv_*
    methods of BiBertEncoder (SimANS/model/models.py:120-160) built on it: values and gradients — including the gradient
    with respect to the supplied embeddings — against the oracle."""
    import torch.nn.functional as F
    from simxns_b200.models import BiBertEncoder
    from test_gpu_parity import build

    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=64)
    pq, pc = bn.synth_params(cfg, seed=31), bn.synth_params(cfg, seed=32)
    nq, nneg, S = 3, 2, 24
    q_ids, q_mask = bn.synth_tokens(nq, 12, cfg["vocab"], seed=1, ragged=True)
    a_ids, a_mask = bn.synth_tokens(nq, S, cfg["vocab"], seed=2, ragged=True, min_len=6)
    b_ids, b_mask = bn.synth_tokens(nq * nneg, S, cfg["vocab"], seed=3, ragged=True, min_len=6)
    rng = np.random.RandomState(4)
    b_emb = (pc["embeddings.word_embeddings.weight"][b_ids] + 0.01 * rng.standard_normal((nq * nneg, S, 128))).astype(np.float32)

    model = BiBertEncoder(None, question_model=build(cfg, pq, X3), ctx_model=build(cfg, pc, X3))
    # plain encode with inputs_embeds == word rows reproduces the id path bit for bit
    ctx = model.ctx_model
    with torch.no_grad():
        via_ids, _ = ctx.encode(cu(b_ids), cu(b_mask))
        via_emb, _ = ctx.encode(None, cu(b_mask), inputs_embeds=ctx.embeddings.word_embeddings.weight[cu(b_ids)])
    assert torch.equal(via_ids, via_emb)

    # oracle for the pair loss: -log_softmax([pos, negs])[:, 1]
    oq, oa, ob = bn.EncoderOracle(cfg, pq), bn.EncoderOracle(cfg, pc), bn.EncoderOracle(cfg, pc)
    q = oq.forward(q_ids, q_mask)["cls"][:, -1]
    a = oa.forward(a_ids, a_mask)["cls"][:, -1]
    b = ob.forward(None, b_mask, inputs_embeds=b_emb)["cls"][:, -1]
    bb = b.reshape(nq, nneg, -1)
    logits = np.concatenate([(q * a).sum(-1, keepdims=True), np.einsum("bh,bdh->bd", q, bb)], axis=1)
    z = logits - logits.max(1, keepdims=True)
    lsm = z - np.log(np.exp(z).sum(1, keepdims=True))
    loss_ref = float(-lsm[:, 1].mean())
    dlog = np.exp(lsm)
    dlog[:, 1] -= 1.0
    dlog /= nq
    dq = dlog[:, :1] * a + np.einsum("bd,bdh->bh", dlog[:, 1:], bb)
    da = dlog[:, :1] * q
    db = (dlog[:, 1:, None] * q[:, None, :]).reshape(nq * nneg, -1)

    def dcls(d):
        out = np.zeros((d.shape[0], 3, 128))
        out[:, -1] = d
        return out

    gq = oq.backward(d_cls=dcls(dq))
    ga = oa.backward(d_cls=dcls(da))
    gb = ob.backward(d_cls=dcls(db))
    gc = {k: ga[k] + gb[k] for k in ga}

    emb_t = cu(b_emb).requires_grad_(True)
    loss, lsm_t = model.forward_adv_pairloss(cu(q_ids), cu(q_mask), cu(a_ids), cu(a_mask), emb_t, cu(b_mask))
    assert abs(loss.item() - loss_ref) < 1e-5
    assert max_rel(lsm_t.detach().cpu().numpy(), lsm) < 1e-5
    loss.backward()
    # the last LayerNorm bias of each tower receives sum_b of softmax-minus-onehot rows: exactly zero
    last_ln = ("encoder.layer.1.output.LayerNorm.bias",)
    assert_grads(model.question_model, gq, 2e-4, zero_true=last_ln)
    assert_grads(model.ctx_model, gc, 2e-4, zero_true=last_ln)
    assert max_rel(emb_t.grad.cpu().numpy(), ob.d_inputs_embeds) < 2e-4
    # padded positions of the supplied embeddings still receive the LayerNorm-path gradient of their own (masked) token:
    # it must be finite and, for tokens no query attends to, independent of the loss only through zero cotangent rows
    assert torch.isfinite(emb_t.grad).all()

    # the two sibling methods: same embeddings, their own loss heads (checked against torch on the kernel's embeddings)
    with torch.no_grad():
        (l3,) = model.forward_adv_triplet(cu(q_ids), cu(q_mask), cu(a_ids), cu(a_mask), cu(b_emb[::nneg].copy()), cu(b_mask[::nneg].copy()))
        l2, zero = model.forward_adv_pairloss_mse(cu(q_ids), cu(q_mask), cu(a_ids), cu(a_mask), cu(b_emb), cu(b_mask))
    b1 = ob.forward(None, b_mask[::nneg], inputs_embeds=b_emb[::nneg])["cls"][:, -1]
    lm = np.stack([(q * a).sum(-1), (q * b1).sum(-1)], 1)
    lm = lm - lm.max(1, keepdims=True)
    want3 = float(-(lm - np.log(np.exp(lm).sum(1, keepdims=True)))[:, 0].mean())
    assert abs(l3.item() - want3) < 1e-5
    want2 = float(((logits[:, 1:] - logits[:, :1]) ** 2).mean())
    assert abs(l2.item() - want2) < 1e-4 * max(1.0, want2) and zero == 0


@pytest.mark.parametrize("GQ,GC,H,q_begin,q_count,c_begin,c_count", [
    (5, 37, 64, 0, 5, 0, 37),              # tiny, nothing aligned to the tile edge
    (130, 1000, 128, 3, 70, 8, 504),       # ranges straddling 64-wide tiles, ragged tails in both dimensions
    (256, 4096, 768, 64, 64, 1024, 1024),  # the 64-wide instantiation, several row tiles
    (1024, 16384, 768, 256, 128, 4096, 2048),  # BASELINE configs[2] full size (128-wide tiles), rank 2 of 8
])
def test_fused_global_nll_launch_vs_unfused_checker_and_oracle(GQ, GC, H, q_begin, q_count, c_begin, c_count):
    """the single cooperative launch (csrc/loss_fused.cu) against (1) the round-1 path kept as its checker (tcgen05 score
    GEMM + row kernel + two gradient GEMMs, SIMX_LOSS_GLOBAL_UNFUSED) and (2) the fp64 oracle: loss, hit count, dQ, dC"""
    import ctypes as C
    rng = np.random.RandomState(GQ + GC)
    Q = (rng.standard_normal((GQ, H)) * (4.0 / np.sqrt(H))).astype(np.float32)
    Cm = (rng.standard_normal((GC, H)) * (4.0 / np.sqrt(H))).astype(np.float32)
    pos = rng.randint(0, GC, size=GQ).astype(np.int32)
    Cm[pos[: GQ // 2]] += 0.5 * Q[: GQ // 2]  # make about half of the positives win their row
    lib = _lib.lib()
    tq, tc, tp = cu(Q), cu(Cm), cu(pos)
    res = {}
    for name, extra in (("fused", 0), ("unfused", _lib.LOSS_GLOBAL_UNFUSED)):
        if extra and c_begin % 8:
            continue  # the checker path needs 8-aligned column offsets (TMA row pitch); the fused kernel does not
        cfg = _lib.LossCfg(terms=_lib.LOSS_NLL_GLOBAL | extra, GQ=GQ, GC=GC, H=H, q_begin=q_begin, q_count=q_count,
                           c_begin=c_begin, c_count=c_count, group=1, w_nll_global=0.7, temperature=1.0, eps=1e-7,
                           score_scale=1.0, grad_scale=2.0)
        ws = torch.empty(lib.simx_score_loss_workspace_bytes(C.byref(cfg)), dtype=torch.uint8, device="cuda")
        out = torch.empty(12, device="cuda")
        dQ, dC = torch.full((q_count, H), 7.0, device="cuda"), torch.full((c_count, H), 7.0, device="cuda")
        _lib.check(lib.simx_score_loss(C.byref(cfg), tq.data_ptr(), tc.data_ptr(), tp.data_ptr(), None, out.data_ptr(),
                                       dQ.data_ptr(), dC.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "simx_score_loss")
        res[name] = (out.cpu().numpy(), dQ.cpu().numpy(), dC.cpu().numpy(), ws.numel())
    # oracle (fp64): full matrices, then this rank's slices
    S = Q.astype(np.float64) @ Cm.astype(np.float64).T
    m = S.max(1, keepdims=True)
    lse = m + np.log(np.exp(S - m).sum(1, keepdims=True))
    nll = float((lse[:, 0] - S[np.arange(GQ), pos]).mean())
    dS = np.exp(S - lse)
    dS[np.arange(GQ), pos] -= 1.0
    dS *= 0.7 * 2.0 / GQ
    want_dQ = dS[q_begin:q_begin + q_count] @ Cm.astype(np.float64)
    want_dC = dS[:, c_begin:c_begin + c_count].T @ Q.astype(np.float64)
    hits = float((S.argmax(1) == pos).sum())
    out, dQ, dC, ws_bytes = res["fused"]
    assert abs(out[1] - nll) < 1e-4 and abs(out[0] - 0.7 * nll) < 1e-4 and out[6] == hits
    assert max_rel(dQ, want_dQ) < 1e-4 and max_rel(dC, want_dC) < 1e-4
    assert ws_bytes < 0.05 * 4 * GQ * GC + 4096, "the fused path must not reserve room for the score matrix"
    if "unfused" in res:
        o2, dQ2, dC2, _ = res["unfused"]
        assert abs(o2[1] - out[1]) < 1e-5 and o2[6] == out[6]
        assert max_rel(dQ, dQ2) < 1e-4 and max_rel(dC, dC2) < 1e-4


def test_lead_distilbert_add_linear_appends_projected_layer():
    """--add_linear (LEAD/models.py:31-77): `linear(hidden -> 768)` on the last hidden state becomes the pooled output and one
    more entry of the per-layer stack; gradients reach the projection and the tower"""
    from types import SimpleNamespace
    cfg = SimpleNamespace(model_type="distilbert", n_layers=2, dim=768, n_heads=12, hidden_dim=1536, vocab_size=300,
                          max_position_embeddings=64, dropout=0.0, attention_dropout=0.0, add_linear=True)
    torch.manual_seed(0)
    enc = lead_models.HFDistilBertEncoder(cfg, precision=X3).cuda()
    ids, mask = bn.synth_tokens(4, 20, 300, seed=2, ragged=True)
    seq, pooled, _, _, layers = enc(input_ids=cu(ids), attention_mask=cu(mask), _want_sequence=True)
    assert tuple(layers.shape) == (4, 3, 1, 768) and tuple(pooled.shape) == (4, 768) and tuple(seq.shape) == (4, 20, 768)
    assert torch.equal(layers[:, -1, 0], pooled)
    raw = enc.last_cls_layers if hasattr(enc, "last_cls_layers") else None
    want = torch.nn.functional.linear(layers[:, 1, 0], enc.linear.weight, enc.linear.bias)  # layer index 1 = last transformer layer
    assert torch.allclose(pooled, want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(seq[:, 0], pooled, rtol=1e-4, atol=1e-5)
    (pooled.square().sum() + layers[:, 0].sum()).backward()
    assert enc.linear.weight.grad.abs().max() > 0 and enc.transformer.layer[0].ffn.lin1.weight.grad.abs().max() > 0
