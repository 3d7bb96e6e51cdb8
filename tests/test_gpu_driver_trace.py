"""B200 (`-m gpu`): the driver-level check.  tests/golden/driver_trace.npz was produced by the reference's OWN `train()`
(SimANS/co_training/co_training_marco_train.py:82-306, imported unmodified and run on the CPU by oracle/make_driver_trace.py
with the reference's model classes, its `get_optimizer`, its schedulers and its accumulate / clip / step tail; only the I/O
— dataset, tensorboard, checkpoint writes — and the missing `transformers.AdamW` are shimmed, outside the package).  Here the
same batches and initial weights go through the swapped-in classes and `simxns_b200.train`, in the loop shape of the driver:
the per-update `loss` / `distill_loss` the driver logs, its learning-rate schedule, and every parameter after the 8th update
must agree (fp32-class mode; Adam turns round-off on near-zero gradients into lr-sized steps, so parameters are judged
against how far the reference itself moved them)."""
import types

import numpy as np
import pytest
import torch

from oracle import bert_numpy as bn
from simxns_b200 import _lib
from simxns_b200.encoder import EncoderConfig
from simxns_b200.models import BiBertEncoder, HFBertEncoder, Reranker
from simxns_b200.train import get_linear_schedule_with_warmup, get_optimizer, retriever_step

pytestmark = pytest.mark.gpu


def test_swapped_import_reproduces_the_reference_train_loop(golden_dir):
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    g = np.load(golden_dir / "driver_trace.npz")
    L, H, heads, I, V, P = (int(x) for x in g["cfg"])
    lr, t_lr, eps, wd, T, max_norm, gas, max_steps, it_step, it_rr = g["hparams"]
    gas, max_steps = int(gas), int(max_steps)
    cfg = bn.bert_config(num_layers=L, hidden=H, heads=heads, intermediate=I, vocab=V, max_pos=P)
    ec = EncoderConfig(num_layers=L, hidden=H, heads=heads, intermediate=I, vocab=V, max_pos=P)
    seeds = [int(s) for s in g["seeds"]]
    pq, pc, pt = (bn.synth_params(cfg, seed=s) for s in seeds[:3])

    def tower(params):
        m = HFBertEncoder(ec, precision=_lib.PREC_BF16X3)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        return m

    model = BiBertEncoder(None, question_model=tower(pq), ctx_model=tower(pc)).cuda()
    teacher = Reranker(tower(pt), H).cuda()
    with torch.no_grad():
        teacher.qa_classifier.weight.copy_(torch.from_numpy(g["head_w"]).cuda())
        teacher.qa_classifier.bias.zero_()
    init = {n: p.detach().clone() for n, p in model.named_parameters()}

    args = types.SimpleNamespace(temperature_distill=float(T), scale_simmila=False, gradient_accumulation_steps=gas,
                                 max_grad_norm=float(max_norm), optimizer="adamW")
    optimizer = get_optimizer(args, model, weight_decay=float(wd), lr=float(lr), eps=float(eps))
    student_max_step = max_steps * (1 - it_rr / it_step)          # co_training_marco_train.py:126-131
    scheduler = get_linear_schedule_with_warmup(optimizer, 0.1 * student_max_step, student_max_step)

    n_batches = sum(1 for k in g.files if k.endswith("/q_ids"))
    t = lambda a: torch.from_numpy(a)
    batches = [dict(student=[t(g[f"b{i}/q_ids"]), t(g[f"b{i}/q_mask"]), t(g[f"b{i}/p_ids"]), t(g[f"b{i}/p_mask"])],
                    teacher=[t(g[f"b{i}/ce_ids"]), t(g[f"b{i}/ce_mask"]), t(g[f"b{i}/tgt"])]) for i in range(n_batches)]

    # the driver's loop shape (:186-263): micro-step counter from 1, the loader restarts when exhausted, an update whenever
    # (step + 1) % gradient_accumulation_steps == 0, logs after every update (logging_steps = 1)
    global_step, step, tr_loss, tr_distill = 0, 0, 0.0, 0.0
    logs = []
    while global_step < max_steps:
        batch = batches[step % n_batches]
        step += 1
        out = retriever_step(args, model, teacher, batch, optimizer, scheduler, step)
        tr_loss += out["loss"].item()
        tr_distill += out["distill_loss"].item()
        if out["grad_norm"] is not None:
            global_step += 1
            logs.append((global_step, tr_loss, tr_distill, scheduler.get_last_lr()[0]))
            tr_loss = tr_distill = 0.0
    assert [l[0] for l in logs] == list(g["log_step"])
    for (k, loss, distill, lr_now), want_loss, want_distill, want_lr in zip(logs, g["log_loss"], g["log_distill"], g["log_lr"]):
        assert abs(lr_now - want_lr) < 1e-12, (k, lr_now, want_lr)
        assert abs(loss - want_loss) < 2e-4 * max(1.0, abs(want_loss)) + 2e-5, (k, loss, want_loss)
        assert abs(distill - want_distill) < 2e-4 * max(1.0, abs(want_distill)) + 4e-5, (k, distill, want_distill)

    worst = ("", 0.0)
    for n, p in model.named_parameters():
        ref = torch.from_numpy(g["student/" + n]).cuda()
        moved = (ref - init[n]).abs().mean().item()
        if "key.bias" in n or "pooler" in n or moved == 0.0 or n.endswith(f"layer.{L - 1}.output.LayerNorm.bias"):
            continue  # true-zero gradients (softmax shift invariance; rows of p_s - p_t sum to zero at the last LayerNorm
            #           bias): Adam turns fp32 round-off into +-lr steps in the reference as well
        d = (p.detach() - ref).abs().mean().item() / moved
        if d > worst[1]:
            worst = (n, d)
    assert worst[1] < 5e-2, worst
    # the teacher only scored (train_flag stayed 0 below iteration_step): untouched, as in the reference
    for n, p in teacher.named_parameters():
        assert torch.equal(p.detach().cpu(), torch.from_numpy(g["teacher/" + n])), n
