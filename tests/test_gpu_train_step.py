"""B200 (`-m gpu`): the co-training loop body (simxns_b200.train: retriever_step / teacher_step — fused KL loss, fused
clip + transformers-flavoured AdamW, linear warm-up schedule, gradient accumulation) against the same loop written with
the reference's own torch lines for the loss (co_training_marco_train.py:199-214, 231-233) and the optimizer oracle
(oracle/optim_numpy.py) for clip_grad_norm_ / AdamW.  fp32-class precision: parameters after the updates within 2e-3 of the step size on average (5e-2 worst element)."""
import copy
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bert_numpy as bn
from oracle import optim_numpy as on
from simxns_b200 import _lib
from simxns_b200.encoder import EncoderConfig
from simxns_b200.models import BiBertEncoder, HFBertEncoder, Reranker
from simxns_b200.train import get_linear_schedule_with_warmup, get_optimizer, retriever_step, teacher_step

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    _lib.require_device()


EC = EncoderConfig(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=400, max_pos=64)
ARGS = types.SimpleNamespace(temperature_distill=2.0, scale_simmila=False, gradient_accumulation_steps=2, max_grad_norm=0.5,
                             optimizer="adamW")
LR, EPS, WD, WARM, TOTAL = 1e-3, 1e-8, 0.01, 1, 4


def make_batches(n, nq=3, group=4):
    out = []
    for i in range(n):
        q_ids, q_mask = bn.synth_tokens(nq, 12, EC.vocab, seed=10 + i, ragged=True)
        p_ids, p_mask = bn.synth_tokens(nq * group, 24, EC.vocab, seed=50 + i, ragged=True, min_len=6)
        ce_ids, ce_mask = bn.synth_tokens(nq * group, 32, EC.vocab, seed=90 + i, ragged=True, min_len=8)
        t = lambda a: torch.from_numpy(a)
        scores = torch.from_numpy((2 * np.random.RandomState(i).standard_normal((nq, group))).astype(np.float32))
        out.append(dict(student=[t(q_ids), t(q_mask), t(p_ids), t(p_mask)],
                        teacher=[t(ce_ids).view(nq, group, 32), t(ce_mask).view(nq, group, 32),
                                 torch.zeros(nq, group, dtype=torch.long)],  # [2] = tgt_tensor of the collate layout
                        teacher_scores=scores))
    return out


def oracle_tail(model, state, step_no, lr):
    """clip_grad_norm_ + transformers.AdamW through the numpy oracle, written back into the model"""
    named = list(model.named_parameters())
    grads = [p.grad.detach().cpu().double().numpy() for _, p in named]
    _, clipped = on.clip_grad_norm(grads, ARGS.max_grad_norm)
    for (n, p), g in zip(named, clipped):
        m, v = state.setdefault(n, (np.zeros(g.shape), np.zeros(g.shape)))
        wd = 0.0 if ("bias" in n or "LayerNorm.weight" in n) else WD
        new_p, m, v = on.adamw_hf(p.detach().cpu().double().numpy(), g, m, v, step_no, lr, eps=EPS, weight_decay=wd)
        state[n] = (m, v)
        p.data.copy_(torch.from_numpy(new_p).to(p.dtype))  # through .data, as transformers.AdamW does: no _version bump
    model.zero_grad()
    for m in model.modules():  # this hand-written update is not a torch Optimizer: tell the towers their weights moved
        if hasattr(m, "invalidate_weights"):
            m.invalidate_weights()


def lr_at(k):  # get_linear_schedule_with_warmup after k scheduler steps
    return LR * (k / max(1, WARM) if k < WARM else max(0.0, (TOTAL - k) / max(1, TOTAL - WARM)))


def assert_same_params(a, b):
    """Adam normalises the gradient: an element whose gradient is ~0 moves by lr * (m / sqrt(v)) where fp32 round-off in g
    is amplified to O(1), so element-wise parity is bounded by a fraction of lr; the BULK of the elements (mean |diff|)
    must agree to 1e-3 of the step size — a wrong bias correction, decay order or clip factor shifts every element by
    O(0.1 lr) and fails both."""
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        # tensors whose TRUE gradient is zero (key biases: softmax shift invariance; the last LayerNorm bias of the tower
        # that feeds a softmax-row loss and the reranker head bias: rows of dS sum to zero) receive pure fp32 round-off,
        # which Adam turns into +-lr steps in the reference as well — nothing to compare
        if "key.bias" in n or n.endswith(f"layer.{EC.num_layers - 1}.output.LayerNorm.bias") or n == "qa_classifier.bias":
            continue
        d = (p - q).abs()
        assert d.max().item() < 0.05 * LR, (n, d.max().item())
        assert d.mean().item() < 2e-3 * LR, (n, d.mean().item())


def test_retriever_steps_match_reference_loop_with_oracle_optimizer():
    torch.manual_seed(0)
    model = BiBertEncoder(None, question_model=HFBertEncoder(EC, precision=_lib.PREC_BF16X3),
                          ctx_model=HFBertEncoder(EC, precision=_lib.PREC_BF16X3)).cuda()
    ref = copy.deepcopy(model)
    opt = get_optimizer(ARGS, model, weight_decay=WD, lr=LR, eps=EPS)
    sched = get_linear_schedule_with_warmup(opt, WARM, TOTAL)
    assert [len(g["params"]) for g in opt.param_groups] == [2 * (3 + 6 * 2 + 1), 2 * (2 + 10 * 2 + 1)]  # decay / no-decay
    state, updates = {}, 0
    for step, batch in enumerate(make_batches(5), 1):
        out = retriever_step(ARGS, model, None, batch, opt, sched, step)
        # the reference's lines (co_training_marco_train.py:198-222) on the twin model
        ref.train()
        q, c = ref(*[t.cuda() for t in batch["student"]])
        sim = torch.einsum("bh,bdh->bd", q, c.reshape(q.size(0), c.size(0) // q.size(0), -1))
        p_s = F.softmax(sim, dim=1)
        p_t = F.softmax(batch["teacher_scores"].cuda() / ARGS.temperature_distill, dim=1)
        distill = torch.nn.KLDivLoss(reduction="batchmean")((p_s + 1e-7).log(), p_t)
        (distill / ARGS.gradient_accumulation_steps).backward()
        assert abs(out["distill_loss"].item() - distill.item()) < 1e-4
        if (step + 1) % ARGS.gradient_accumulation_steps == 0:  # the reference's test (:246): micro-steps 1, 3, 5
            updates += 1
            oracle_tail(ref, state, updates, lr_at(updates - 1))
            assert out["grad_norm"] is not None and out["grad_norm"].item() > ARGS.max_grad_norm  # the clip was active
            assert_same_params(model, ref)
            assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in model.parameters())
        else:
            assert out["grad_norm"] is None
    assert updates == 3 and abs(opt.param_groups[0]["lr"] - lr_at(3)) < 1e-12


def test_teacher_steps_match_reference_loop_with_oracle_optimizer():
    torch.manual_seed(1)
    args = copy.copy(ARGS)
    args.gradient_accumulation_steps = 1
    teacher = Reranker(HFBertEncoder(EC, precision=_lib.PREC_BF16X3), EC.hidden).cuda()
    ref = copy.deepcopy(teacher)
    opt = get_optimizer(args, teacher, weight_decay=WD, lr=LR, eps=EPS)
    sched = get_linear_schedule_with_warmup(opt, WARM, TOTAL)
    state = {}
    for step, batch in enumerate(make_batches(3), 1):
        out = teacher_step(args, teacher, batch, opt, sched, step)
        ref.train()
        logits = ref(batch["teacher"][0].cuda(), batch["teacher"][1].cuda())
        loss = torch.nn.CrossEntropyLoss()(logits, torch.zeros(logits.size(0), dtype=torch.long, device="cuda"))
        loss.backward()
        assert abs(out["contr_loss"].item() - loss.item()) < 1e-4
        oracle_tail(ref, state, step, lr_at(step - 1))
        assert_same_params(teacher, ref)


def test_data_style_optimizer_updates_reach_the_kernels():
    """an optimizer that writes through `p.data` (transformers.AdamW v4 does) never bumps `p._version`; the towers' bf16
    operand cache must still see its update (global optimizer-step hook)"""

    class DataSGD(torch.optim.Optimizer):
        def __init__(self, params):
            super().__init__(params, dict(lr=0.5))

        def step(self, closure=None):
            for g in self.param_groups:
                for p in g["params"]:
                    if p.grad is not None:
                        p.data.add_(p.grad.data, alpha=-g["lr"])

    torch.manual_seed(2)
    enc = HFBertEncoder(EC, precision=_lib.PREC_BF16X3).cuda()
    ids, mask = bn.synth_tokens(3, 16, EC.vocab, seed=1)
    ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    opt = DataSGD(enc.parameters())
    _, pooled, _ = enc(input_ids=ids, attention_mask=mask)
    pooled.square().sum().backward()
    v0 = enc.embeddings.word_embeddings.weight._version
    opt.step()
    assert enc.embeddings.word_embeddings.weight._version == v0  # the hazard: nothing visible on the tensor
    with torch.no_grad():
        _, after, _ = enc(input_ids=ids, attention_mask=mask)
    fresh = HFBertEncoder(EC, precision=_lib.PREC_BF16X3).cuda()
    fresh.load_state_dict(enc.state_dict())
    with torch.no_grad():
        _, want, _ = fresh(input_ids=ids, attention_mask=mask)
    assert (after - pooled.detach()).abs().max().item() > 1e-3  # the step moved the output ...
    assert torch.equal(after, want)                              # ... exactly to what the new weights give


def test_two_stream_tower_overlap_changes_nothing_but_the_schedule():
    """BiBertEncoder issues the query tower on a second CUDA stream (models.BiBertEncoder._towers).  Same kernels, same
    inputs: embeddings bit-identical, gradients identical up to the order of the split-K / column-sum atomics; weight
    updates made on the main stream between two calls are seen by the side stream; and outputs consumed on the main stream
    right away (loss, backward) are complete.  (Several optimizer steps are NOT compared: Adam turns the atomics' rounding
    noise on true-zero gradients into +-lr updates, so two runs of the SAME schedule drift apart as well.)"""
    from simxns_b200.losses import simans_distill_loss
    from test_gpu_parity import BF16, build, cu

    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=400, max_pos=160)
    pq, pc = bn.synth_params(cfg, seed=3), bn.synth_params(cfg, seed=4)
    q_ids, q_mask = bn.synth_tokens(16, 32, cfg["vocab"], seed=5, ragged=True)
    p_ids, p_mask = bn.synth_tokens(64, 128, cfg["vocab"], seed=6, ragged=True)
    teacher = torch.randn(16, 4, generator=torch.Generator().manual_seed(7)).cuda()
    outs = []
    for overlap in (False, True):
        model = BiBertEncoder(None, question_model=build(cfg, pq, BF16), ctx_model=build(cfg, pc, BF16)).train()
        model.overlap_towers = overlap
        q, a = model(cu(q_ids), cu(q_mask), cu(p_ids), cu(p_mask))
        loss = simans_distill_loss(q, a, teacher, temperature_distill=1.0)
        loss.backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        with torch.no_grad():  # an in-place update of every master weight, issued on the main stream
            for p in model.parameters():
                p.mul_(1.01)
            q2, a2 = model(cu(q_ids), cu(q_mask), cu(p_ids), cu(p_mask))
        outs.append((loss.item(), q.detach().clone(), a.detach().clone(), grads, q2.clone(), a2.clone()))
    (l0, q0, a0, g0, q20, a20), (l1, q1, a1, g1, q21, a21) = outs
    assert torch.equal(q0, q1) and torch.equal(a0, a1) and abs(l0 - l1) < 1e-6
    assert torch.equal(q20, q21) and torch.equal(a20, a21) and not torch.equal(q20, q0)
    gmax = max(float(g.abs().max()) for g in g0.values())
    for n in g0:
        err = float((g0[n] - g1[n]).abs().max())
        assert err <= 2e-3 * max(float(g0[n].abs().max()), 1e-2 * gmax), (n, err)
