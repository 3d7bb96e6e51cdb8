"""Score / loss / gradient step on libsimx_b200 (`simx_score_loss`) behind the reference's loss interfaces.

  fused_score_loss(...)            the general entry: any sum of
                                     nll_global  in-batch InfoNCE over all-gathered embeddings  (BiEncoderNllLoss.calc,
                                                 /root/reference/SimANS/model/models.py:468-505; gathered as in
                                                 /root/reference/MASTER/finetune/MS/run_de_model.py:180-234 and
                                                 /root/reference/LEAD/util.py:200-204,416-451)
                                     kl_local    SimANS KLDiv(log(softmax(s)+eps) || softmax(teacher/T))
                                                 (/root/reference/SimANS/co_training/co_training_marco_train.py:199-214)
                                     ce_local    -sum p_t log(p_s+eps)/B (wiki / MASTER / PROD prob_loss)
                                     kd_local    PROD KD_softmax (/root/reference/PROD/ProD_KD/model/models.py:772-777)
                                     nll_local   PROD hard loss, positive = column 0 (.../models.py:694-712)
                                     mse_local   PROD KD_logit, 0.5 MSE(scores, teacher logits) (.../models.py:778-780)
                                     dkd_local   PROD DKD, alpha TCKD + beta NCKD (.../models.py:783-812)
  BiEncoderNllLoss / CrossBERTKDLoss / simans_distill_loss   drop-in call signatures of the reference.

The kernel returns the loss terms AND d(loss)/d(local embeddings) in the same pass; autograd only scales them by the
incoming cotangent.  Cross-rank negatives: the local [CLS] embeddings are all-gathered with NCCL over NVLink
(`dist.all_gather_into_tensor`), remote slices are constants, gradients reach the local rows/columns only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib

_TERMS = {"nll_global": _lib.LOSS_NLL_GLOBAL, "kl_local": _lib.LOSS_KL_LOCAL, "ce_local": _lib.LOSS_CE_LOCAL,
          "kd_local": _lib.LOSS_KD_LOCAL, "nll_local": _lib.LOSS_NLL_LOCAL, "adv_local": _lib.LOSS_ADV_LOCAL,
          "mse_local": _lib.LOSS_MSE_LOCAL, "dkd_local": _lib.LOSS_DKD_LOCAL}
LOSS_FIELDS = ("total", "nll_global", "kl_local", "ce_local", "kd_local", "nll_local", "n_correct", "n_correct_local",
               "mse_local", "dkd_local")


def gather_embeddings(local: torch.Tensor, group=None) -> torch.Tensor:
    """[n, H] -> [world*n, H] (rank-major), every rank contributing the same n.  No autograd: gradients for the local
    slice come from the fused kernel (the reference re-inserts the local tensor for the same effect, LEAD/util.py:200-204)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class _ScoreLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_local, c_local, teacher, pos_idx, opts):
        lib = _lib.lib()
        dev = q_local.device
        if dev.type != "cuda":
            raise RuntimeError("fused_score_loss needs CUDA tensors (no CPU fallback)")
        with torch.cuda.device(dev):
            _lib.require_device()
            q_local = q_local.detach().to(torch.float32).contiguous()
            c_local = c_local.detach().to(torch.float32).contiguous()
            nq, H = q_local.shape
            nc = c_local.shape[0]
            terms = 0
            for t in opts["terms"]:
                terms |= _TERMS[t]
            use_global = bool(terms & _lib.LOSS_NLL_GLOBAL)
            if opts.get("unfused_check"):
                terms |= _lib.LOSS_GLOBAL_UNFUSED
            rank, world = 0, 1
            if use_global and opts.get("gather", True) and dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(opts.get("group")), dist.get_world_size(opts.get("group"))
            if world > 1:
                Q = gather_embeddings(q_local, opts.get("group"))
                Cm = gather_embeddings(c_local, opts.get("group"))
            else:
                Q, Cm = q_local, c_local
            GQ, GC = Q.shape[0], Cm.shape[0]
            group = opts.get("group_size") or (nc // nq)
            w = opts.get("weights", {})
            cfg = _lib.LossCfg(
                terms=terms, GQ=GQ, GC=GC, H=H, q_begin=rank * nq, q_count=nq, c_begin=rank * nc, c_count=nc,
                group=group, w_nll_global=w.get("nll_global", 1.0), w_kl_local=w.get("kl_local", 1.0),
                w_ce_local=w.get("ce_local", 1.0), w_kd_local=w.get("kd_local", 1.0), w_nll_local=w.get("nll_local", 1.0),
                temperature=opts.get("temperature", 1.0), eps=opts.get("eps", 1e-7),
                score_scale=opts.get("score_scale", 1.0), grad_scale=1.0, w_adv_local=w.get("adv_local", 1.0),
                w_mse_local=w.get("mse_local", 1.0), w_dkd_local=w.get("dkd_local", 1.0),
                dkd_alpha=opts.get("dkd_alpha", 1.0), dkd_beta=opts.get("dkd_beta", 1.0))
            if use_global:
                if pos_idx is None:
                    pos_idx = torch.arange(0, GQ * group, group, device=dev, dtype=torch.int32)  # LEAD/util.py:445
                pos_idx = pos_idx.to(device=dev, dtype=torch.int32).contiguous()
                if pos_idx.numel() != GQ:
                    raise ValueError(f"pos_idx has {pos_idx.numel()} entries for {GQ} global queries")
            if teacher is not None:
                teacher = teacher.detach().to(device=dev, dtype=torch.float32).contiguous()
                if tuple(teacher.shape) != (nq, group):
                    raise ValueError(f"teacher logits must be [{nq}, {group}], got {tuple(teacher.shape)}")
            ws_bytes = lib.simx_score_loss_workspace_bytes(C.byref(cfg))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            losses = torch.empty(12, dtype=torch.float32, device=dev)
            dQ = torch.empty_like(q_local)
            dC = torch.empty_like(c_local)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.simx_score_loss(C.byref(cfg), Q.data_ptr(), Cm.data_ptr(), _lib.ptr(pos_idx), _lib.ptr(teacher),
                                           losses.data_ptr(), dQ.data_ptr(), dC.data_ptr(), ws.data_ptr(), stream),
                       "simx_score_loss")
        ctx.save_for_backward(dQ, dC)
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        dQ, dC = ctx.saved_tensors
        return dQ * g_total, dC * g_total, None, None, None


def fused_score_loss(q_local: torch.Tensor, c_local: torch.Tensor, terms: Sequence[str], *, teacher=None, pos_idx=None,
                     weights: Optional[dict] = None, temperature: float = 1.0, eps: float = 1e-7,
                     score_scale: float = 1.0, group_size: Optional[int] = None, gather: bool = True, group=None,
                     dkd_alpha: float = 1.0, dkd_beta: float = 1.0, unfused_check: bool = False):
    """-> (loss scalar with autograd into q_local / c_local, detail tensor [12] = LOSS_FIELDS + pad).
    unfused_check: evaluate the global term through the 3-GEMM checker path instead of the fused launch."""
    opts = dict(terms=tuple(terms), weights=dict(weights or {}), temperature=float(temperature), eps=float(eps),
                score_scale=float(score_scale), group_size=group_size, gather=gather, group=group,
                dkd_alpha=float(dkd_alpha), dkd_beta=float(dkd_beta), unfused_check=bool(unfused_check))
    return _ScoreLossFn.apply(q_local, c_local, teacher, pos_idx, opts)


# ---------------------------------------------------------------------------------------------------------
# reference-shaped wrappers
# ---------------------------------------------------------------------------------------------------------
class BiEncoderNllLoss(object):
    """SimANS/model/models.py:468-514 (identical class in MASTER / PROD / CAPSTONE).  q_vectors / ctx_vectors are the
    ALREADY-GATHERED global embeddings in the reference's call sites (run_de_model.py:180-234), so no gather here."""

    def calc(self, q_vectors, ctx_vectors, positive_idx_per_question: list, hard_negative_idx_per_question: list = None,
             loss_scale: float = None):
        pos = torch.as_tensor(positive_idx_per_question, dtype=torch.int32, device=q_vectors.device)
        loss, detail = fused_score_loss(q_vectors, ctx_vectors, ["nll_global"], pos_idx=pos, gather=False,
                                        group_size=max(1, ctx_vectors.shape[0] // q_vectors.shape[0]))
        correct_predictions_count = detail[6].round().to(torch.int64)
        if loss_scale:
            loss = loss * loss_scale
        return loss, correct_predictions_count

    @staticmethod
    def get_scores(q_vector, ctx_vectors):
        return torch.matmul(q_vector, torch.transpose(ctx_vectors, 0, 1))

    @staticmethod
    def get_similarity_function():
        return BiEncoderNllLoss.get_scores


class CrossBERTKDLoss(object):
    """PROD/ProD_KD/model/models.py:668-781: CE_WEIGHT * NLL(column 0) + KD_WEIGHT * soft [+ LwF_WEIGHT * LwF] on the
    query's own [1+n] block.  Every KD_type of the reference ('KD_softmax', 'KD_logit', 'DKD', 'prob_loss') runs fused;
    LwF adds a second fused call against the frozen student's scores (kd_loss handles KD_softmax / KD_logit only there,
    as in the reference)."""

    _SOFT = {"KD_softmax": "kd_local", "KD_logit": "mse_local", "DKD": "dkd_local", "prob_loss": "ce_local"}


    def calc(self, args, q_vectors, ctx_vectors, relevance_logits, hard_negative_idx_per_question: list = None,
             loss_scale: float = None, LwF=False, ori_q_vector=None, ori_ctx_vectors=None):
        nq = q_vectors.shape[0]
        group = ctx_vectors.shape[0] // nq
        if relevance_logits is None:
            # no teacher: the bare hard loss — weighted by CE_WEIGHT only on the LwF branch (models.py:768-774)
            loss, detail = fused_score_loss(q_vectors, ctx_vectors, ["nll_local"], group_size=group, gather=False,
                                            weights={"nll_local": args.CE_WEIGHT if LwF else 1.0})
        elif args.KD_type in self._SOFT:
            soft = self._SOFT[args.KD_type]
            alpha, beta = getattr(args, "DKD_alpha", None), getattr(args, "DKD_beta", None)
            if args.KD_type == "DKD" and (alpha is None or beta is None):
                raise ValueError("DKD loss need refer DKD_alpha and DKD_beta in args")  # the reference prints this and exits
            loss, detail = fused_score_loss(q_vectors, ctx_vectors, ["nll_local", soft], teacher=relevance_logits,
                                            weights={"nll_local": args.CE_WEIGHT, soft: args.KD_WEIGHT},
                                            temperature=args.TEMPERATURE, group_size=group, gather=False,
                                            dkd_alpha=1.0 if alpha is None else alpha,
                                            dkd_beta=1.0 if beta is None else beta)
        else:
            raise ValueError("no such type of KD loss,please check KD_type")
        if LwF:
            if args.KD_type not in ("KD_softmax", "KD_logit"):  # kd_loss has no other branch (UnboundLocalError there)
                raise ValueError(f"LwF needs KD_type KD_softmax or KD_logit, got {args.KD_type!r}")
            with torch.no_grad():  # the frozen copy's scores are targets only (its gradients feed no optimizer)
                ori = torch.einsum("bh,bdh->bd", ori_q_vector.float(), ori_ctx_vectors.float().reshape(nq, group, -1))
            lwf, _ = fused_score_loss(q_vectors, ctx_vectors, [self._SOFT[args.KD_type]], teacher=ori,
                                      temperature=args.TEMPERATURE, group_size=group, gather=False)
            loss = loss + args.LwF_WEIGHT * lwf
        correct_predictions_count = detail[7].round().to(torch.int64)  # local rows whose positive (column 0) ranks first
        if loss_scale:
            loss = loss * loss_scale
        return loss, correct_predictions_count


def simans_wiki_loss(local_q_vector, local_ctx_vectors, relevance_logits, temperature_normal: float = 1.0,
                     adv_lambda: float = 0.0, eps: float = 1e-7, scale_simmila: bool = False):
    """The NQ/TQ retriever step's loss (SimANS/wiki/co_training_wiki_train.py:198-228):
    adv_lambda * sum(reward * log(p_s+eps)) + (1 - adv_lambda) * (-sum p_t log(p_s+eps) / B), one fused kernel call."""
    nq = local_q_vector.shape[0]
    scale = 1.0 / (local_q_vector.shape[1] ** 0.5) if scale_simmila else 1.0
    terms = ["ce_local"] + (["adv_local"] if adv_lambda != 0 else [])
    loss, _ = fused_score_loss(local_q_vector, local_ctx_vectors, terms, teacher=relevance_logits,
                               weights=dict(ce_local=1.0 - adv_lambda, adv_local=adv_lambda),
                               temperature=temperature_normal, eps=eps, score_scale=scale,
                               group_size=local_ctx_vectors.shape[0] // nq, gather=False)
    return loss


def simans_distill_loss(local_q_vector, local_ctx_vectors, relevance_logits, temperature_distill: float = 1.0,
                        eps: float = 1e-7, scale_simmila: bool = False):
    """The retriever step's loss, co_training_marco_train.py:199-214, as one fused kernel call."""
    nq = local_q_vector.shape[0]
    scale = 1.0 / (local_q_vector.shape[1] ** 0.5) if scale_simmila else 1.0
    loss, _ = fused_score_loss(local_q_vector, local_ctx_vectors, ["kl_local"], teacher=relevance_logits,
                               temperature=temperature_distill, eps=eps, score_scale=scale,
                               group_size=local_ctx_vectors.shape[0] // nq, gather=False)
    return loss


# ---------------------------------------------------------------------------------------------------------
# LEAD: mutual distillation between two dual encoders on the global score matrices (simx_pair_distill)
# ---------------------------------------------------------------------------------------------------------
PAIR_FIELDS = ("total", "nll_a", "nll_b", "kl_b2a", "kl_a2b")


class _PairDistillFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_a, c_a, q_b, c_b, opts):
        lib = _lib.lib()
        dev = q_a.device
        if dev.type != "cuda":
            raise RuntimeError("fused_pair_distill needs CUDA tensors (no CPU fallback)")
        with torch.cuda.device(dev):
            _lib.require_device()
            loc = [t.detach().to(torch.float32).contiguous() for t in (q_a, c_a, q_b, c_b)]
            nq, H = loc[0].shape
            nc = loc[1].shape[0]
            rank, world = 0, 1
            if opts.get("gather", True) and dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(opts.get("group")), dist.get_world_size(opts.get("group"))
            glob = [gather_embeddings(t, opts.get("group")) for t in loc] if world > 1 else loc
            GQ, GC = glob[0].shape[0], glob[1].shape[0]
            group = opts.get("group_size") or (nc // nq)
            pos = opts.get("pos_idx")
            if pos is None:
                pos = torch.arange(0, GQ * group, group, device=dev, dtype=torch.int32)  # LEAD/util.py:596
            pos = torch.as_tensor(pos, device=dev).to(torch.int32).contiguous()
            if pos.numel() != GQ:
                raise ValueError(f"pos_idx has {pos.numel()} entries for {GQ} global queries")
            w = opts["weights"]
            train_b = any(ctx.needs_input_grad[2:4])
            cfg = _lib.PairCfg(GQ=GQ, GC=GC, H=H, q_begin=rank * nq, q_count=nq, c_begin=rank * nc, c_count=nc,
                               w_nll_a=w.get("nll_a", 0.0), w_nll_b=w.get("nll_b", 0.0), w_kl_b2a=w.get("kl_b2a", 0.0),
                               w_kl_a2b=w.get("kl_a2b", 0.0), temperature=opts.get("temperature", 1.0), grad_scale=1.0,
                               w_kd_b2a=w.get("kd_b2a", 0.0), w_mse_b2a=w.get("mse_b2a", 0.0),
                               w_dkd_b2a=w.get("dkd_b2a", 0.0), dkd_alpha=opts.get("dkd_alpha", 1.0),
                               dkd_beta=opts.get("dkd_beta", 1.0))
            ws = torch.empty(lib.simx_pair_distill_workspace_bytes(GQ, GC, H), dtype=torch.uint8, device=dev)
            losses = torch.empty(8, dtype=torch.float32, device=dev)
            grads = [torch.empty_like(loc[0]), torch.empty_like(loc[1]),
                     torch.empty_like(loc[2]) if train_b else None, torch.empty_like(loc[3]) if train_b else None]
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.simx_pair_distill(C.byref(cfg), glob[0].data_ptr(), glob[1].data_ptr(), glob[2].data_ptr(),
                                             glob[3].data_ptr(), pos.data_ptr(), losses.data_ptr(), grads[0].data_ptr(),
                                             grads[1].data_ptr(), _lib.ptr(grads[2]), _lib.ptr(grads[3]), ws.data_ptr(),
                                             stream), "simx_pair_distill")
        ctx.grads = grads
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g):
        return tuple(None if g is None else g * g_total for g in ctx.grads) + (None,)


def fused_pair_distill(q_a, c_a, q_b, c_b, weights: dict, *, temperature: float = 1.0, group_size: Optional[int] = None,
                       gather: bool = True, group=None, pos_idx=None, dkd_alpha: float = 1.0, dkd_beta: float = 1.0):
    """a = student, b = teacher (local embeddings, all-gathered inside).  weights: nll_a, nll_b, kl_b2a (teacher targets ->
    student), kl_a2b.  -> (total with autograd into all four inputs, detail [8] = PAIR_FIELDS + pad)."""
    opts = dict(weights=dict(weights), temperature=float(temperature), group_size=group_size, gather=gather, group=group,
                pos_idx=pos_idx, dkd_alpha=float(dkd_alpha), dkd_beta=float(dkd_beta))
    return _PairDistillFn.apply(q_a, c_a, q_b, c_b, opts)


class _EmbedMseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.lib()
        if a.device.type != "cuda":
            raise RuntimeError("embed_mse needs CUDA tensors (no CPU fallback)")
        with torch.cuda.device(a.device):
            _lib.require_device()
            a32 = a.detach().to(torch.float32).contiguous()
            b32 = b.detach().to(device=a.device, dtype=torch.float32).contiguous()
            if a32.shape != b32.shape:
                raise ValueError(f"embed_mse: shapes differ {tuple(a32.shape)} vs {tuple(b32.shape)}")
            loss = torch.zeros((), dtype=torch.float32, device=a.device)
            grad = torch.empty_like(a32)
            _lib.check(lib.simx_embed_mse(a32.data_ptr(), b32.data_ptr(), a32.numel(), 1.0, loss.data_ptr(), grad.data_ptr(),
                                          torch.cuda.current_stream(a.device).cuda_stream), "simx_embed_mse")
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        return ctx.grad * g, None


def embed_mse(student: torch.Tensor, teacher: torch.Tensor) -> torch.Tensor:
    """0.5 * MSELoss(mean)(student, teacher) on raw embeddings, gradient into `student` only (one half of bi_logit)"""
    return _EmbedMseFn.apply(student, teacher)


class BiEncoderKDLoss(object):
    """PROD/ProD_KD/model/models.py:970-1038: DE -> DE distillation on the (already gathered) global matrix,
    CE_WEIGHT * NLL + KD_WEIGHT * soft, soft = KD_softmax (T^2 KL(softmax(teacher/T) || softmax(student/T))), KD_logit
    (0.5 MSE of the score matrices), DKD, or Bi_logit (0.5 MSE of the raw embeddings, models.py:1062-1067)."""

    def calc(self, args, q_vectors, ctx_vectors, teacher_q_vector, teacher_ctxs_vector, positive_idx_per_question: list,
             hard_negative_idx_per_question: list = None, loss_scale: float = None):
        if teacher_q_vector is None or teacher_ctxs_vector is None:
            return BiEncoderNllLoss().calc(q_vectors, ctx_vectors, positive_idx_per_question, loss_scale=loss_scale)
        if args.KD_type == "Bi_logit":
            loss, correct_predictions_count = BiEncoderNllLoss().calc(q_vectors, ctx_vectors, positive_idx_per_question)
            soft = embed_mse(q_vectors, teacher_q_vector) + embed_mse(ctx_vectors, teacher_ctxs_vector)
            loss = args.CE_WEIGHT * loss + args.KD_WEIGHT * soft
            if loss_scale:
                loss = loss * loss_scale
            return loss, correct_predictions_count
        key = {"KD_softmax": "kd_b2a", "KD_logit": "mse_b2a", "DKD": "dkd_b2a"}.get(args.KD_type)
        if key is None:
            raise ValueError("no such type of KD loss,please check KD_type")
        alpha, beta = getattr(args, "DKD_alpha", None), getattr(args, "DKD_beta", None)
        if args.KD_type == "DKD" and (alpha is None or beta is None):
            raise ValueError("DKD loss need refer DKD_alpha and DKD_beta in args")
        loss, detail = fused_pair_distill(q_vectors, ctx_vectors, teacher_q_vector.detach(), teacher_ctxs_vector.detach(),
                                          {"nll_a": args.CE_WEIGHT, key: args.KD_WEIGHT}, temperature=args.TEMPERATURE,
                                          gather=False, pos_idx=positive_idx_per_question,
                                          dkd_alpha=1.0 if alpha is None else alpha,
                                          dkd_beta=1.0 if beta is None else beta)
        correct_predictions_count = detail[6].round().to(torch.int64)
        if loss_scale:
            loss = loss * loss_scale
        return loss, correct_predictions_count


def lead_distill_loss(q_db, c_db, q_de, c_de, cls_q_db=None, cls_c_db=None, cls_q_de=None, cls_c_de=None, *,
                      selected_db: Sequence[int] = (), selected_de: Sequence[int] = (), temperature: float = 1.0,
                      layer_temperature: float = 1.0, para_db: float = 1.0, para_de: float = 1.0, para_dis: float = 1.0,
                      para_layer: float = 1.0, reweight: bool = True, train_de: bool = True, gather: bool = True):
    """DE-teacher -> DB-student branch of LEAD's distill_loss with layer_score_dis_loss
    (/root/reference/LEAD/util.py:539-559, 594-644, 737-770).

    q_*/c_*: local final embeddings; cls_*: per-layer [CLS] rows `[n, L, H]` in LEAD's layer indexing (layer k =
    `cls_layers[:, k+1]` of the encoder, i.e. pass `encoder.last_cls_layers[:, 1:]`).  Only the [CLS] rows are gathered —
    the reference all-gathers the full [n, L, S, H] hidden states and then reads [:, l, 0, :] (util.py:572-573, 605).
    Returns (loss, dict of detached component values)."""
    w = dict(nll_a=para_db, nll_b=para_de if train_de else 0.0, kl_b2a=para_dis, kl_a2b=para_dis if train_de else 0.0)
    loss, detail = fused_pair_distill(q_db, c_db, q_de, c_de, w, temperature=temperature, gather=gather)
    parts = dict(loss_db=detail[1], loss_dual=detail[2], loss_de_db_dis=detail[3], loss_db_de_dis=detail[4])
    if len(selected_db):
        kls, nlls = [], []
        for ls, lt in zip(selected_db, selected_de):
            kl, d = fused_pair_distill(cls_q_db[:, ls], cls_c_db[:, ls], cls_q_de[:, lt], cls_c_de[:, lt],
                                       dict(kl_b2a=1.0), temperature=temperature, gather=gather)
            kls.append(kl)
            nlls.append(d[2])
        if reweight:  # lambda = softmax(-NLL(teacher layer scores) / T_layer), a constant (util.py:548-550)
            lam = torch.softmax(-torch.stack(nlls).detach() / layer_temperature, dim=0)
        else:
            lam = torch.full((len(kls),), float(para_layer), device=loss.device)
        layer_loss = sum(l * k for l, k in zip(lam, kls))
        loss = loss + layer_loss
        parts["loss_de_db_layer_dis"] = layer_loss.detach()
    return loss, parts


class _LocalPairFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, c, teacher, opts):
        lib = _lib.lib()
        dev = q.device
        if dev.type != "cuda":
            raise RuntimeError("fused_local_pair_distill needs CUDA tensors (no CPU fallback)")
        with torch.cuda.device(dev):
            _lib.require_device()
            q32, c32 = q.detach().to(torch.float32).contiguous(), c.detach().to(torch.float32).contiguous()
            t32 = teacher.detach().to(device=dev, dtype=torch.float32).contiguous()
            nq, H = q32.shape
            nc = c32.shape[0]
            if nc % nq or tuple(t32.shape) != (nq, nc // nq):
                raise ValueError(f"teacher logits must be [{nq}, {nc // max(nq, 1)}], got {tuple(t32.shape)}")
            w = opts["weights"]
            cfg = _lib.PairCfg(GQ=nq, GC=nc, H=H, q_begin=0, q_count=nq, c_begin=0, c_count=nc, w_nll_a=w.get("nll_a", 0.0),
                               w_nll_b=w.get("nll_b", 0.0), w_kl_b2a=w.get("kl_b2a", 0.0), w_kl_a2b=w.get("kl_a2b", 0.0),
                               temperature=opts["temperature"], grad_scale=1.0, w_kd_b2a=0.0)
            losses = torch.empty(8, dtype=torch.float32, device=dev)
            dq, dc = torch.empty_like(q32), torch.empty_like(c32)
            dt = torch.empty_like(t32) if ctx.needs_input_grad[2] else None
            _lib.check(lib.simx_local_pair_distill(C.byref(cfg), q32.data_ptr(), c32.data_ptr(), t32.data_ptr(),
                                                   losses.data_ptr(), dq.data_ptr(), dc.data_ptr(), _lib.ptr(dt),
                                                   torch.cuda.current_stream(dev).cuda_stream), "simx_local_pair_distill")
        ctx.grads = (dq, dc, dt)
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g):
        return tuple(None if g is None else g * g_total for g in ctx.grads) + (None,)


def fused_local_pair_distill(q_local, c_local, teacher_logits, weights: dict, *, temperature: float = 1.0):
    """Student dual encoder (a) vs cross-encoder relevance logits (b) on each query's own [1+n] block.
    weights: nll_a, nll_b, kl_b2a, kl_a2b.  Gradients flow into q_local, c_local and (if it requires grad) teacher_logits."""
    return _LocalPairFn.apply(q_local, c_local, teacher_logits, dict(weights=dict(weights), temperature=float(temperature)))


def lead_ce_distill_loss(q_db, c_db, ce_logits, cls_q_db=None, cls_c_db=None, ce_layer_logits=None, *,
                         selected_db: Sequence[int] = (), selected_ce: Sequence[int] = (), temperature: float = 1.0,
                         layer_temperature: float = 1.0, para_db: float = 1.0, para_ce: float = 1.0, para_dis: float = 1.0,
                         train_ce: bool = True):
    """CE-teacher -> DB-student branch of LEAD's distill_loss (/root/reference/LEAD/util.py:669-717, 801-829), no
    all-gather (util.py:567).  ce_logits [B, 1+n]; ce_layer_logits [L_t, B, 1+n] (third output of lead_models.Reranker);
    cls_*_db: per-layer [CLS] rows [n, L, H] of the student."""
    w = dict(nll_a=para_db, nll_b=para_ce if train_ce else 0.0, kl_b2a=para_dis, kl_a2b=para_dis if train_ce else 0.0)
    loss, detail = fused_local_pair_distill(q_db, c_db, ce_logits, w, temperature=temperature)
    parts = dict(loss_db=detail[1], loss_ce=detail[2], loss_ce_db_dis=detail[3], loss_db_ce_dis=detail[4])
    if len(selected_db):
        kls, nlls = [], []
        for ls, lt in zip(selected_db, selected_ce):
            kl, d = fused_local_pair_distill(cls_q_db[:, ls], cls_c_db[:, ls], ce_layer_logits[lt], dict(kl_b2a=1.0),
                                             temperature=temperature)
            kls.append(kl)
            nlls.append(d[2])
        lam = torch.softmax(-torch.stack(nlls).detach() / layer_temperature, dim=0)
        layer_loss = sum(l * k for l, k in zip(lam, kls))
        loss = loss + layer_loss
        parts["loss_ce_db_layer_dis"] = layer_loss.detach()
    return loss, parts


def gather_tensor(local: torch.Tensor, group=None) -> torch.Tensor:
    """LEAD/util.py:200-204 `gather_tensor`: all-gather along dim 0 with the LOCAL slice re-inserted, so gradients reach the
    local rows only and remote rows are constants.  One NCCL all-gather into a flat buffer instead of a list of tensors."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = local.shape[0]
    out = local.new_empty((world * n,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.detach().contiguous(), group=group)
    return torch.cat([out[:rank * n], local, out[(rank + 1) * n:]], dim=0)


class _ColbertScoreFn(torch.autograd.Function):
    """late-interaction scores in document chunks: the reference materialises the whole [Q*Sq, D*Sd] token-pair matrix
    (4.3 GB fp32 at 128 queries x 2048 passages); here a chunk of it lives at a time and backward rebuilds the one-hot
    routing of each chunk from the saved arg-max positions"""

    @staticmethod
    def forward(ctx, q, d, keep, chunk, fill=-9e9):
        Q, Sq, H = q.shape
        D, Sd, _ = d.shape
        qf = q.reshape(Q * Sq, H)
        out = q.new_empty(Q, D)
        idx = torch.empty(Q, Sq, D, dtype=torch.int64, device=q.device)
        for lo in range(0, D, chunk):
            hi = min(lo + chunk, D)
            logits = torch.matmul(qf, d[lo:hi].reshape(-1, H).t()).view(Q, Sq, hi - lo, Sd)
            logits.masked_fill_(~keep[lo:hi][None, None], fill)
            mx, ix = logits.max(3)
            out[:, lo:hi] = mx.sum(1)
            idx[:, :, lo:hi] = ix
        ctx.save_for_backward(q, d, keep, idx)
        ctx.chunk = chunk
        return out

    @staticmethod
    def backward(ctx, g):
        q, d, keep, idx = ctx.saved_tensors
        Q, Sq, H = q.shape
        D, Sd, _ = d.shape
        qf = q.reshape(Q * Sq, H)
        dq = torch.zeros_like(qf)
        dd = torch.empty_like(d)
        for lo in range(0, D, ctx.chunk):
            hi = min(lo + ctx.chunk, D)
            w = q.new_zeros(Q, Sq, hi - lo, Sd)
            w.scatter_(3, idx[:, :, lo:hi, None], g[:, None, lo:hi, None].expand(Q, Sq, hi - lo, 1))
            w.mul_(keep[lo:hi][None, None])  # a filled position passes no gradient (index_fill_ in the reference)
            wf = w.view(Q * Sq, (hi - lo) * Sd)
            dq.addmm_(wf, d[lo:hi].reshape(-1, H))
            dd[lo:hi] = torch.matmul(wf.t(), qf).view(hi - lo, Sd, H)
        return dq.view(Q, Sq, H), dd, None, None, None


def colbert_score(q_embs_col, d_embs_col, attention_mask, q_num, chunk: int = 256, fill: float = -9e9):
    """LEAD/util.py:459-465: MaxSim late interaction.  q_embs_col [Q, Sq, H], d_embs_col [D, Sd, H], attention_mask [D, Sd]
    of the documents -> scores [q_num, D * Q / q_num]: sum over ALL query positions (padded ones included, as in the
    reference) of the maximum over the un-masked document tokens of q_i . d_j."""
    keep = attention_mask.reshape(d_embs_col.shape[0], d_embs_col.shape[1]) != 0
    scores = _ColbertScoreFn.apply(q_embs_col, d_embs_col, keep, int(chunk), float(fill))
    return scores.view(q_num, -1)


def get_loss_cross(args, relevance_logits):
    """LEAD/util.py:410-414: cross-entropy of the cross-encoder's [B, 1+n] relevance logits, positive in column 0"""
    target = torch.zeros(relevance_logits.size(0), dtype=torch.long, device=relevance_logits.device)
    return torch.nn.functional.cross_entropy(relevance_logits, target)


def get_loss_dual(args, local_q_vector, local_ctx_vectors, attention_mask, reduction='mean'):
    """LEAD/util.py:416-451 (run_single_model.py warm-up of every tower type): in-batch InfoNCE over the embeddings of all
    ranks, positives at `i * (1+n)`.  Dual encoders: the fused global kernel on the all-gathered [CLS] embeddings.
    `'colbert' in args.model_type`: late-interaction scores (masked document tokens filled with -1e4, as there) on the
    all-gathered token embeddings.  The reference gathers the embeddings but not `attention_mask`; here the mask travels
    with them."""
    if reduction != 'mean':
        raise NotImplementedError("get_loss_dual: only reduction='mean' (the only value the reference's callers pass)")
    if 'colbert' in getattr(args, "model_type", ""):
        q, c, m = gather_tensor(local_q_vector.contiguous()), gather_tensor(local_ctx_vectors.contiguous()), gather_tensor(attention_mask)
        q_num = q.size(0)
        scores = colbert_score(q, c, m, q_num, fill=-1e4)
        target = torch.arange(0, c.size(0), c.size(0) // q_num, device=scores.device, dtype=torch.long)
        return torch.nn.functional.nll_loss(torch.nn.functional.log_softmax(scores, dim=1), target)
    group = local_ctx_vectors.shape[0] // local_q_vector.shape[0]
    loss, _ = fused_score_loss(local_q_vector, local_ctx_vectors, ["nll_global"], group_size=group, gather=True)
    return loss


def lead_col_distill_loss(q_db, c_db, q_col, c_col, doc_mask, cls_q_db=None, cls_c_db=None, hid_q_col=None, hid_c_col=None, *,
                          selected_db: Sequence[int] = (), selected_col: Sequence[int] = (), temperature: float = 1.0,
                          layer_temperature: float = 1.0, para_db: float = 1.0, para_col: float = 1.0, para_dis: float = 1.0,
                          para_layer: float = 1.0, reweight: bool = True, train_col: bool = True, gather: bool = True):
    """ColBERT-teacher -> DB-student branch of LEAD's distill_loss (/root/reference/LEAD/util.py:581-585 gather, :594-615
    student scores, :647-664 teacher scores, :730-735 + :772-794 loss sum, :539-559 layer re-weighting) — the two shipped
    scripts distillation/distill_from_12cb_to_6de.sh and distill_from_24cb_to_12de.sh.

    q_db / c_db: the student's final embeddings [n, H]; q_col / c_col: the teacher's normalised token embeddings
    [n, S, 768] (HFColBertEncoder output 0); doc_mask [n_docs, S_d]; cls_*_db [n, L_s, H] the student's per-layer [CLS]
    rows; hid_*_col [n, L_t, S, H] the teacher's per-layer hidden states (HFColBertEncoder output 1).  The score matrices
    are small ([GQ, GC]); the arithmetic on them is the reference's, term by term, with NO detach (targets carry gradient
    to whichever model produced them).  Only the selected teacher layers are gathered.  Returns (loss, dict of parts)."""
    import torch.nn.functional as F

    g = gather_tensor if gather else (lambda t: t)
    q_db, c_db, q_col, c_col, doc_mask = g(q_db), g(c_db), g(q_col), g(c_col), g(doc_mask)
    q_num = q_db.shape[0]
    sample_num = c_db.shape[0] // q_num
    target = torch.arange(0, q_num * sample_num, sample_num, device=q_db.device, dtype=torch.long)
    T = float(temperature)

    def views(scores):
        return F.log_softmax(scores, dim=-1), F.softmax(scores / T, dim=-1)

    lsm_db, tgt_db = views(torch.matmul(q_db, c_db.t()).view(q_num, -1))
    lsm_col, tgt_col = views(colbert_score(q_col, c_col, doc_mask, q_num))
    parts = {}
    loss_db = F.nll_loss(lsm_db, target)
    loss = para_db * loss_db
    parts["loss_db"] = loss_db.detach()
    if train_col:
        loss_col = F.nll_loss(lsm_col, target)
        loss = loss + para_col * loss_col
        parts["loss_col"] = loss_col.detach()
    loss_col_db_dis = F.kl_div(lsm_db, tgt_col, reduction="batchmean")
    loss_db_col_dis = F.kl_div(lsm_col, tgt_db, reduction="batchmean")
    loss = loss + para_dis * ((loss_col_db_dis + loss_db_col_dis) if train_col else loss_col_db_dis)
    parts["loss_col_db_dis"], parts["loss_db_col_dis"] = loss_col_db_dis.detach(), loss_db_col_dis.detach()
    if len(selected_db):
        kls, lams = [], []
        for ls, lt in zip(selected_db, selected_col):
            lsm_s, _ = views(torch.matmul(g(cls_q_db[:, ls]), g(cls_c_db[:, ls]).t()).view(q_num, -1))
            lsm_t, tgt_t = views(colbert_score(g(hid_q_col[:, lt]), g(hid_c_col[:, lt]), doc_mask, q_num))
            kls.append(F.kl_div(lsm_s, tgt_t, reduction="batchmean"))
            lams.append(-F.nll_loss(lsm_t, target).detach() / layer_temperature)  # torch.tensor([...]) in the reference: detached
        lam = torch.softmax(torch.stack(lams), dim=0) if reweight else torch.full((len(kls),), float(para_layer),
                                                                                  device=loss.device)
        layer_loss = sum(l * k for l, k in zip(lam, kls))
        loss = loss + layer_loss
        parts["loss_col_db_layer_dis"] = layer_loss.detach()
    return loss, parts


def query_doc_attention_map(args, q_all_layer_hidden, d_all_layer_hidden, query_len, doc_len, selected_index_list):
    """LEAD/util.py:504-529 (VIRT): per selected layer the token-level map softmax_over_doc_tokens(q_i . d_j) for every
    (query, document) pair, positions beyond query_len[i] / doc_len[j] filled with -1e9 before the softmax.
    q_all_layer_hidden [Q, L, Sq, H], d_all_layer_hidden [D, L, Sd, H] (towers with `full_layer_hidden = True`) ->
    (maps, log-maps, maps at args.temperature), each a list of [Q, D, Sq, Sd].  Same values as the reference; the mask is
    built from two comparisons instead of a Python loop over Q x D."""
    import torch.nn.functional as F

    dev = q_all_layer_hidden.device
    Sq, Sd = q_all_layer_hidden.shape[2], d_all_layer_hidden.shape[2]
    qv = torch.arange(Sq, device=dev)[None, :] < torch.as_tensor(query_len, device=dev)[:, None]   # [Q, Sq]
    dv = torch.arange(Sd, device=dev)[None, :] < torch.as_tensor(doc_len, device=dev)[:, None]     # [D, Sd]
    masked = ~(qv[:, None, :, None] & dv[None, :, None, :])                                          # [Q, D, Sq, Sd]
    maps, scores, targets = [], [], []
    for l in selected_index_list:
        att = torch.einsum("ijk,mnk->imjn", q_all_layer_hidden[:, l], d_all_layer_hidden[:, l])     # [Q, D, Sq, Sd]
        att = att.masked_fill(masked, -1e9)
        scores.append(F.log_softmax(att, dim=-1))
        targets.append(F.softmax(att / args.temperature, dim=-1))
        maps.append(F.softmax(att, dim=-1))
    return maps, scores, targets


def virt_loss(query_doc_attention_t, query_doc_attention_s, distill_para):
    """LEAD/util.py:531-537: sum of squared differences of the teacher's and the student's maps, averaged over layers,
    queries and documents (NOT over token positions), times distill_para"""
    total = 0
    for t, s_ in zip(query_doc_attention_t, query_doc_attention_s):
        total = total + ((s_ - t) ** 2).sum()
    return total / len(query_doc_attention_t) / query_doc_attention_t[0].shape[0] / query_doc_attention_t[0].shape[1] * distill_para


def select_layer(args):
    """LEAD's choice of the layer pairs whose [CLS] scores are distilled (/root/reference/LEAD/util.py:206-254), re-drawn
    every `save_steps` by run_LEAD.py.  Returns (selected_index_list_db, selected_index_list_teacher) for
    lead_distill_loss / lead_ce_distill_loss.  `layer_selection_random` consumes Python's global `random` stream exactly
    as the reference does (student list first, then the teacher list)."""
    from random import sample

    selected_index_list_db, selected_index_list_teacher = None, None
    if args.layer_selection_random:
        large = 'large' in args.pretrained_model_name
        extra = 1 if args.add_linear else 0
        n_db, n_teacher = (12, 24) if large else (6, 12)
        selected_index_list_db = sorted(sample([i for i in range(n_db + extra)], args.disitll_layer_num))
        selected_index_list_teacher = sorted(sample([i for i in range(n_teacher + extra)], args.disitll_layer_num))
    elif args.layer_selection_last:
        if 'ms' in args.train_file:  # passage distillation: 5 layers, document distillation: 6
            selected_index_list_db, selected_index_list_teacher = [1, 2, 3, 4, 5], [7, 8, 9, 10, 11]
        else:
            selected_index_list_db, selected_index_list_teacher = [0, 1, 2, 3, 4, 5], [6, 7, 8, 9, 10, 11]
    elif args.layer_selection_skip:
        if 'ms' in args.train_file:
            selected_index_list_db, selected_index_list_teacher = [1, 2, 3, 4, 5], [0, 2, 4, 6, 8]
        else:
            selected_index_list_db, selected_index_list_teacher = [0, 1, 2, 3, 4, 5], [0, 2, 4, 6, 8, 10]
    return selected_index_list_db, selected_index_list_teacher


def global_positive_indices(local_positive_idxs, n_local_ctx: int, device=None, group=None) -> torch.Tensor:
    """positives of ALL ranks' queries as columns of the rank-major gathered passage matrix: rank r's local index v
    becomes v + (number of passages of ranks < r) — the bookkeeping of `caculate_cont_loss`
    (/root/reference/MASTER/finetune/MS/run_de_model.py:203-220; same loop in PROD run_progressive_distill_marco.py:371-447).
    Every rank must contribute the same number of queries and passages (the gathered matrices are dense)."""
    idx = torch.as_tensor(list(local_positive_idxs), dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return idx.to(torch.int32)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = idx.new_empty(world * idx.numel())
    dist.all_gather_into_tensor(out, idx + rank * int(n_local_ctx), group=group)
    return out.to(torch.int32)


def caculate_cont_loss(args, local_q_vector, local_ctx_vectors, local_positive_idxs):
    """Drop-in for `caculate_cont_loss` (run_de_model.py:180-234, spelling as in the reference): in-batch InfoNCE over the
    embeddings of every rank.  The reference pickles the vectors, all-reduces a 640 MB byte buffer and rebuilds the
    matrices on every rank; here the [CLS] rows are all-gathered over NCCL inside the fused loss and only the positive
    indices travel separately.  -> (loss, is_correct)"""
    pos = global_positive_indices(local_positive_idxs, local_ctx_vectors.shape[0], device=local_q_vector.device)
    loss, detail = fused_score_loss(local_q_vector, local_ctx_vectors, ["nll_global"], pos_idx=pos, gather=True,
                                    group_size=max(1, local_ctx_vectors.shape[0] // local_q_vector.shape[0]))
    return loss, detail[6].round().to(torch.int64)
