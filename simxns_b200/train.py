"""The body of the SimANS co-training loop as callable steps on the B200 path
(/root/reference/SimANS/co_training/co_training_marco_train.py: `get_optimizer` :57-69, the `train_flag == 0` retriever
step :195-224, the `train_flag == 1` teacher step :225-245, and the accumulate / clip / step / zero_grad tail :246-263).

The reference has no `train_step` function — the step is the inline body of `train()`; these functions keep its argument
names (`args.temperature_distill`, `args.scale_simmila`, `args.gradient_accumulation_steps`, `args.max_grad_norm`) and
its order of operations so that the loop reads

    for step, batch in enumerate(loader, 1):
        out = retriever_step(args, model, teacher_model, batch, optimizer, scheduler, step)      # train_flag == 0
        out = teacher_step(args, teacher_model, batch, teacher_optimizer, teacher_scheduler, step)  # train_flag == 1

Everything heavy runs in libsimx_b200: both towers (tcgen05 GEMMs, attention), the fused KL loss, and the optimizer tail
(one norm launch + one AdamW launch for all parameters).  PyTorch is the host.  Differences that do NOT change the
result: the two `.item()` host syncs per step (:223-224) are replaced by device tensors in the returned dict.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from . import optim as sx_optim
from .losses import simans_distill_loss


def get_optimizer(args, model: nn.Module, weight_decay: float = 0.0, lr: float = 0.0, eps: float = 0.0,
                  flavor: str = "hf", sync=None) -> torch.optim.Optimizer:
    """co_training_marco_train.py:57-69 — two groups, `bias` / `LayerNorm.weight` without decay; transformers.AdamW
    semantics by default (`flavor="torch"` gives LEAD's torch.optim.AdamW, LEAD/util.py:889-901).
    sync: a `gradsync.FlatGradSync(model, shard=True)` wrapper -> the sharded optimizer (optimizer state / update split
    over the ranks)."""
    no_decay = ['bias', 'LayerNorm.weight']
    optimizer_grouped_parameters = [
        {'params': [p for n, p in model.named_parameters() if p.requires_grad and not any(nd in n for nd in no_decay)],
         'weight_decay': weight_decay},
        {'params': [p for n, p in model.named_parameters() if p.requires_grad and any(nd in n for nd in no_decay)],
         'weight_decay': 0.0}
    ]
    if getattr(args, "optimizer", "adamW") != "adamW":
        raise Exception("optimizer {0} not recognized! Can only be adamW".format(args.optimizer))
    if sync is not None:
        return sx_optim.ShardedAdamW(sync, optimizer_grouped_parameters, lr=lr, eps=eps, flavor=flavor)
    return sx_optim.AdamW(optimizer_grouped_parameters, lr=lr, eps=eps, flavor=flavor)


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, last_epoch: int = -1):
    """transformers.get_linear_schedule_with_warmup (co_training_marco_train.py:129-134): linear warm-up to the base lr,
    then linear decay to 0.  A LambdaLR, so it drives `group["lr"]` of the fused optimizer like any other."""

    def lr_lambda(current_step: int):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        return max(0.0, float(num_training_steps - current_step) / float(max(1, num_training_steps - num_warmup_steps)))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda, last_epoch)


def _ids(t, dev):
    """the driver casts every id / mask tensor with .long() (:189-194); int32 batches (data.collate(dtype=torch.int32)) keep
    their width — the kernels read both"""
    return (t if t.dtype == torch.int32 else t.long()).to(dev)


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def _finish(args, model, optimizer, scheduler, step: int) -> Optional[torch.Tensor]:
    """:246-263 — every `gradient_accumulation_steps` micro-steps: clip, optimizer.step, scheduler.step, zero_grad.
    The clip is folded into the fused AdamW launch (the norm never visits the host); returns the gradient norm.
    The reference tests `(step + 1) % gradient_accumulation_steps == 0` with `step` counted from 1 (:192, :246): with the
    shipped MS-Pas setting of 2 (train_MS_Pas_AR2.sh:11) the optimizer steps after micro-steps 1, 3, 5, ... — the first
    update sees a single micro-batch.  Kept as is: same updates as the reference driver."""
    if (step + 1) % args.gradient_accumulation_steps != 0:
        return None
    if isinstance(optimizer, sx_optim.AdamW):
        optimizer.step(max_grad_norm=args.max_grad_norm)
        norm = optimizer.last_grad_norm
    else:
        norm = sx_optim.clip_grad_norm_(_unwrap(model).parameters(), args.max_grad_norm)
        optimizer.step()
    if scheduler is not None:
        scheduler.step()
    model.zero_grad()
    return norm


def retriever_step(args, model, teacher_model, batch, optimizer, scheduler, step: int) -> Dict[str, torch.Tensor]:
    """train_flag == 0 (:195-224): the teacher scores the batch under no_grad, the student is distilled with
    KLDiv(log(softmax(q.c)+eps) || softmax(teacher / T)), backward through both towers, then the optimizer tail."""
    dev = next(_unwrap(model).parameters()).device
    batch_student = batch['student']
    inputs_student = {"query_ids": _ids(batch_student[0], dev), "attention_mask_q": _ids(batch_student[1], dev),
                      "input_ids_a": _ids(batch_student[2], dev), "attention_mask_a": _ids(batch_student[3], dev)}
    model.train()
    local_q_vector, local_ctx_vectors = model(**inputs_student)
    if teacher_model is not None:
        teacher_model.eval()
        with torch.no_grad():
            relevance_logits = teacher_model(input_ids=_ids(batch['teacher'][0], dev), attention_mask=_ids(batch['teacher'][1], dev))
    else:
        # pre-computed teacher scores [B, 1+n] under their own key.  (batch['teacher'][2] of the collate layout is
        # tgt_tensor — the 0/1 positive labels, MARCO_until_new.py:255 — never scores: it is not read here.)
        relevance_logits = batch.get('teacher_scores') if hasattr(batch, "get") else None
        if relevance_logits is None:
            raise ValueError("retriever_step(teacher_model=None) needs batch['teacher_scores'] [B, 1+n] (float logits)")
        if not torch.is_floating_point(relevance_logits):
            raise TypeError(f"batch['teacher_scores'] must be floating-point logits, got {relevance_logits.dtype}")
        relevance_logits = relevance_logits.to(dev)
    distill_loss = simans_distill_loss(local_q_vector, local_ctx_vectors, relevance_logits,
                                       temperature_distill=args.temperature_distill, eps=1e-7,
                                       scale_simmila=getattr(args, "scale_simmila", False))
    loss = distill_loss / args.gradient_accumulation_steps
    loss.backward()
    norm = _finish(args, model, optimizer, scheduler, step)
    return dict(loss=loss.detach(), distill_loss=distill_loss.detach(), grad_norm=norm)


def teacher_step(args, teacher_model, batch, teacher_optimizer, teacher_scheduler, step: int) -> Dict[str, torch.Tensor]:
    """train_flag == 1 (:225-245): cross-encoder CE over the [B, 1+n] relevance logits with the positive in column 0."""
    dev = next(_unwrap(teacher_model).parameters()).device
    teacher_model.train()
    relevance_logits = teacher_model(input_ids=_ids(batch['teacher'][0], dev), attention_mask=_ids(batch['teacher'][1], dev))
    relevance_target = torch.zeros(relevance_logits.size(0), dtype=torch.long, device=dev)
    contr_loss = torch.nn.functional.cross_entropy(relevance_logits, relevance_target)
    loss = contr_loss / args.gradient_accumulation_steps
    loss.backward()
    norm = _finish(args, teacher_model, teacher_optimizer, teacher_scheduler, step)
    return dict(loss=loss.detach(), contr_loss=contr_loss.detach(), grad_norm=norm)
