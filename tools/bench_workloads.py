#!/usr/bin/env python
"""Throughput of the BASELINE.json configs beyond the headline, on 1..8 B200s (one rank per GPU under torchrun, NCCL):

  prod_kd        configs[3]  PROD ProD_KD, cross-encoder teacher -> 6-layer student (PROD/README.md:214-226:
                             --max_seq_length=144 --per_gpu_train_batch_size=8 --number_neg 15 --open_LwF
                             --teacher_num_hidden_layers 12 --student_num_hidden_layers 6 --KD_type=KD_softmax --CE_WEIGHT 0.1
                             --KD_WEIGHT 0.9 --TEMPERATURE 4.0 --LwF_WEIGHT 1.0): teacher scores [q,16,176] under no_grad, the
                             frozen initial student scores the batch for LwF, the student (DDP, dropout 0.1) takes
                             CrossBERTKDLoss (hard NLL + T^2 KL + LwF) forward + backward.
  lead           configs[4]  LEAD 12-layer BERT dual-encoder teacher / 6-layer DistilBERT student, both trained
  lead_col       LEAD 12-layer ColBERT teacher / 6-layer DistilBERT student (distill_from_12cb_to_6de.sh), late interaction
                             (LEAD/distillation/distill_from_12de_to_6de.sh: --share_weight --distill_de --train_de --distill_db
                             --train_db --distill_de_db_layer_score --layer_score_reweight, 5 layer pairs): four tower passes,
                             all-gathered [CLS] rows of 1 + 5 layers, mutual KL + per-layer re-weighted KL, backward through
                             both models (DDP on both).
  teacher_large  N2 at the shipped teacher width: Reranker CE step over ERNIE-large (24 L, H 1024, 16 heads, I 4096;
                             SimANS/train_MS_Pas_AR2.sh:13,23: --gradient_checkpointing) at [q,16,160].

    python tools/bench_workloads.py --workload prod_kd [--queries 8]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_workloads.py --workload lead

Timing: CUDA events on the launching stream between barriers, 3 warm-up steps, max over ranks; one JSON line per
measurement on rank 0 (`--out` appends them to a file under profiles/).  Weak scaling: per-GPU work is fixed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def fwd_flops(layers, S, H=768, I=3072):
    return layers * (8 * S * H * H + 4 * S * H * I + 4 * S * S * H)


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json), sustained bf16"
    return 1400.0, "fallback (B200_PROFILING.md)"


class Runner:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
                os.environ.pop("NCCL_DEBUG")
            dist.init_process_group("nccl", device_id=self.dev)

    def ddp(self, model):
        if self.world == 1:
            return model
        if os.environ.get("SIMX_BENCH_DDP", "0") == "1":
            return self.torch.nn.parallel.DistributedDataParallel(model, device_ids=[self.local], find_unused_parameters=False,
                                                                  gradient_as_bucket_view=True)
        from simxns_b200.gradsync import FlatGradSync

        return FlatGradSync(model)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup=3):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1) / steps
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def tokens(self, n, S, vocab, seed):
        torch = self.torch
        g = torch.Generator().manual_seed(seed + 1000 * self.rank)
        ids = torch.randint(1000, vocab, (n, S), generator=g)
        ids[:, 0] = 101
        return ids.to(self.dev), torch.ones(n, S, dtype=torch.long, device=self.dev)

    def line(self, **kw):
        peak, src = peaks()
        kw.update(n_gpus=self.world, scaling="weak", dtype="bf16", data="synthetic", peak_tflops=peak, peak_source=src)
        if "model_tflops_per_gpu" in kw:
            kw["frac_of_peak"] = kw["model_tflops_per_gpu"] / peak
        kw["peak_mem_gb"] = self.torch.cuda.max_memory_allocated() / 2 ** 30
        if self.rank == 0:
            print(json.dumps(kw), flush=True)
        return kw


def set_dropout(model, p):
    for m in model.modules():
        if hasattr(m, "hidden_dropout_prob"):
            m.hidden_dropout_prob = m.attention_probs_dropout_prob = p


def bench_prod_kd(R: Runner, args):
    torch = R.torch
    from simxns_b200 import _lib
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.prod_models import BiBertEncoder, CrossBERTKDLoss, HFBertEncoder, Reranker

    out = []
    for nq in args.queries:
        torch.manual_seed(0)
        group, q_len, p_len, ce_len = 16, 32, 144, 176
        s_cfg, t_cfg = EncoderConfig(num_layers=6), EncoderConfig(num_layers=12)
        student = BiBertEncoder(None, question_model=HFBertEncoder(s_cfg, precision=_lib.PREC_BF16),
                                ctx_model=HFBertEncoder(s_cfg, precision=_lib.PREC_BF16)).to(R.dev)
        frozen = BiBertEncoder(None, question_model=HFBertEncoder(s_cfg, precision=_lib.PREC_BF16),
                               ctx_model=HFBertEncoder(s_cfg, precision=_lib.PREC_BF16)).to(R.dev).eval()  # LwF copy
        frozen.load_state_dict(student.state_dict())
        teacher = Reranker(HFBertEncoder(t_cfg, precision=_lib.PREC_BF16), 768).to(R.dev).eval()
        set_dropout(student, args.dropout)
        student.train()
        step_model = R.ddp(student)
        q, qm = R.tokens(nq, q_len, s_cfg.vocab, 1)
        p, pm = R.tokens(nq * group, p_len, s_cfg.vocab, 2)
        ce, cem = R.tokens(nq * group, ce_len, s_cfg.vocab, 3)
        ce, cem = ce.view(nq, group, ce_len), cem.view(nq, group, ce_len)
        a = types.SimpleNamespace(KD_type="KD_softmax", CE_WEIGHT=0.1, KD_WEIGHT=0.9, TEMPERATURE=4.0, LwF_WEIGHT=1.0)
        kd = CrossBERTKDLoss()

        def step():
            with torch.no_grad():
                _, relevance_logits, _ = teacher(ce, cem)
                oq, oc = frozen(q, qm, p, pm)
            sq, sc = step_model(q, qm, p, pm)
            loss, _ = kd.calc(a, sq, sc, relevance_logits, LwF=True, ori_q_vector=oq, ori_ctx_vectors=oc)
            loss.backward()
            student.zero_grad(set_to_none=True)
            return loss

        ms = R.timed(step, args.steps)
        pairs = nq * group * R.world / (ms * 1e-3)
        fl = nq * group * fwd_flops(12, ce_len) + 4 * (nq * fwd_flops(6, q_len) + nq * group * fwd_flops(6, p_len))
        out.append(R.line(workload=f"prod_kd (BASELINE configs[3]): 12L cross-encoder teacher [q,16,{ce_len}] no_grad + frozen 6L "
                                   f"student (LwF) + 6L student fwd/bwd, CrossBERTKDLoss KD_softmax T=4 + LwF, {nq} queries/GPU x 16, "
                                   f"q_len {q_len}, p_len {p_len}, dropout {args.dropout}",
                          metric="query+passage pairs/sec", value=pairs, unit="pairs/s", ms_per_step=ms, steps=args.steps,
                          queries_per_gpu=nq, model_tflops_per_gpu=fl / (ms * 1e-3) / 1e12))
        del student, frozen, teacher, step_model
        torch.cuda.empty_cache()
    return out


def bench_lead(R: Runner, args):
    torch = R.torch
    from simxns_b200 import _lib, lead_models
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.losses import lead_distill_loss

    out = []
    for nq in args.queries:
        torch.manual_seed(0)
        group, q_len, p_len = 16, 32, 128
        t_cfg = EncoderConfig(num_layers=12)
        s_cfg = EncoderConfig(arch="distilbert", num_layers=6, type_vocab=0)
        # --share_weight: one tower encodes queries and passages in both models
        de = lead_models.BiBertEncoder(None, question_model=lead_models.HFBertEncoder(t_cfg, precision=_lib.PREC_BF16)).to(R.dev)
        db = lead_models.BiBertEncoder(None, question_model=lead_models.HFDistilBertEncoder(s_cfg, precision=_lib.PREC_BF16)).to(R.dev)
        set_dropout(de, args.dropout)
        set_dropout(db, args.dropout)
        de.train()
        db.train()
        de_step, db_step = R.ddp(de), R.ddp(db)
        q, qm = R.tokens(nq, q_len, t_cfg.vocab, 1)
        p, pm = R.tokens(nq * group, p_len, t_cfg.vocab, 2)
        sel_db, sel_de = [1, 2, 3, 4, 5], [7, 8, 9, 10, 11]  # select_layer(layer_selection_last), 'ms' passage distillation

        def step():
            q_de, c_de, ql_de, cl_de, _ = de_step(q, qm, p, pm)
            q_db, c_db, ql_db, cl_db, _ = db_step(q, qm, p, pm)
            loss, _ = lead_distill_loss(q_db, c_db, q_de, c_de, ql_db[:, :, 0], cl_db[:, :, 0], ql_de[:, :, 0], cl_de[:, :, 0],
                                        selected_db=sel_db, selected_de=sel_de, temperature=1.0, layer_temperature=10.0,
                                        reweight=True, train_de=True, gather=True)
            loss.backward()
            de.zero_grad(set_to_none=True)
            db.zero_grad(set_to_none=True)
            return loss

        ms = R.timed(step, args.steps)
        pairs = nq * group * R.world / (ms * 1e-3)
        fl = 3 * (nq * fwd_flops(12, q_len) + nq * group * fwd_flops(12, p_len)) + \
            3 * (nq * fwd_flops(6, q_len) + nq * group * fwd_flops(6, p_len))
        out.append(R.line(workload=f"lead (BASELINE configs[4]): 12L BERT DE teacher + 6L DistilBERT student, both trained "
                                   f"(share_weight), mutual KL + NLL on all-gathered [CLS] + 5 re-weighted layer pairs, "
                                   f"{nq} queries/GPU x 16, q_len {q_len}, p_len {p_len}, dropout {args.dropout}",
                          metric="query+passage pairs/sec", value=pairs, unit="pairs/s", ms_per_step=ms, steps=args.steps,
                          queries_per_gpu=nq, model_tflops_per_gpu=fl / (ms * 1e-3) / 1e12))
        del de, db, de_step, db_step
        torch.cuda.empty_cache()
    return out


def bench_lead_col(R: Runner, args):
    """LEAD distillation/distill_from_12cb_to_6de.sh: 12-layer ColBERT teacher (trained) -> 6-layer DistilBERT student, late
    interaction on the final token embeddings and on 5 selected teacher layers (full hidden states with gradients)"""
    torch = R.torch
    from simxns_b200 import _lib, lead_models
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.losses import lead_col_distill_loss

    out = []
    for nq in args.queries:
        torch.manual_seed(0)
        group, q_len, p_len = 16, 32, 128
        t_cfg = EncoderConfig(num_layers=12)
        s_cfg = EncoderConfig(arch="distilbert", num_layers=6, type_vocab=0)
        col = lead_models.BiBertEncoder(types.SimpleNamespace(model_type="colbert"),
                                        question_model=lead_models.HFColBertEncoder(t_cfg, precision=_lib.PREC_BF16)).to(R.dev)
        db = lead_models.BiBertEncoder(None, question_model=lead_models.HFDistilBertEncoder(s_cfg, precision=_lib.PREC_BF16)).to(R.dev)
        set_dropout(col, args.dropout)
        set_dropout(db, args.dropout)
        col.train()
        db.train()
        col_step, db_step = R.ddp(col), R.ddp(db)
        q, qm = R.tokens(nq, q_len, t_cfg.vocab, 1)
        p, pm = R.tokens(nq * group, p_len, t_cfg.vocab, 2)
        sel_db, sel_col = [1, 2, 3, 4, 5], [7, 8, 9, 10, 11]

        def step():
            q_col, c_col, qh_col, ch_col, _ = col_step(q, qm, p, pm)
            q_db, c_db, ql_db, cl_db, _ = db_step(q, qm, p, pm)
            loss, _ = lead_col_distill_loss(q_db, c_db, q_col, c_col, pm, ql_db[:, :, 0], cl_db[:, :, 0], qh_col, ch_col,
                                            selected_db=sel_db, selected_col=sel_col, temperature=1.0, layer_temperature=10.0,
                                            reweight=True, train_col=True, gather=True)
            loss.backward()
            col.zero_grad(set_to_none=True)
            db.zero_grad(set_to_none=True)
            return loss

        ms = R.timed(step, args.steps)
        pairs = nq * group * R.world / (ms * 1e-3)
        gq, gd = nq * R.world, nq * group * R.world
        late = 6 * 3 * 2.0 * gq * q_len * gd * p_len * 768           # 6 score matrices, forward + two backward products
        fl = 3 * (nq * fwd_flops(12, q_len) + nq * group * fwd_flops(12, p_len)) + \
            3 * (nq * fwd_flops(6, q_len) + nq * group * fwd_flops(6, p_len)) + late
        out.append(R.line(workload=f"lead_col (distill_from_12cb_to_6de.sh): 12L ColBERT teacher + 6L DistilBERT student, both "
                                   f"trained (share_weight), MaxSim on final + 5 teacher layers (full hidden states), "
                                   f"{nq} queries/GPU x 16, q_len {q_len}, p_len {p_len}, dropout {args.dropout}",
                          metric="query+passage pairs/sec", value=pairs, unit="pairs/s", ms_per_step=ms, steps=args.steps,
                          queries_per_gpu=nq, model_tflops_per_gpu=fl / (ms * 1e-3) / 1e12,
                          late_interaction_tflop_per_step=late / 1e12))
        del col, db, col_step, db_step
        torch.cuda.empty_cache()
    return out


def bench_teacher_large(R: Runner, args):
    torch = R.torch
    from simxns_b200 import _lib
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.models import HFBertEncoder, Reranker

    out = []
    ec = EncoderConfig(num_layers=24, hidden=1024, heads=16, intermediate=4096)
    for nq in args.queries:
        for ck in (True, False):
            T = nq * 16 * 160
            saved_gb = T * 24 * 16 * 1024 * 2 / 2 ** 30
            if not ck and saved_gb > 120:
                continue  # the full activation set does not fit beside weights / gradients / workspace
            torch.manual_seed(0)
            torch.cuda.reset_peak_memory_stats()
            rr = Reranker(HFBertEncoder(ec, precision=_lib.PREC_BF16), 1024).to(R.dev).train()
            set_dropout(rr, args.dropout)
            if ck:
                rr.encoder.gradient_checkpointing_enable()
            step_model = R.ddp(rr)
            ids, mask = R.tokens(nq * 16, 160, ec.vocab, 5)
            ids, mask = ids.view(nq, 16, 160), mask.view(nq, 16, 160)
            tgt = torch.zeros(nq, dtype=torch.long, device=R.dev)

            def step():
                loss = torch.nn.functional.cross_entropy(step_model(ids, mask), tgt)  # co_training_marco_train.py:231-233
                loss.backward()
                rr.zero_grad(set_to_none=True)

            ms = R.timed(step, args.steps)
            pairs = nq * 16 * R.world / (ms * 1e-3)
            fl = 3 * nq * 16 * fwd_flops(24, 160, 1024, 4096)
            out.append(R.line(workload=f"teacher_large (N2 at the shipped width): Reranker CE step, ERNIE-large shape 24L/H1024/16 heads/"
                                       f"I4096, [{nq}, 16, 160], dropout {args.dropout}, "
                                       f"{'gradient checkpointing (layer inputs only)' if ck else 'all activations saved'}",
                              metric="query+passage pairs/sec", value=pairs, unit="pairs/s", ms_per_step=ms, steps=args.steps,
                              queries_per_gpu=nq, gradient_checkpointing=ck, model_tflops_per_gpu=fl / (ms * 1e-3) / 1e12,
                              executed_tflops_per_gpu=(4 if ck else 3) / 3 * fl / (ms * 1e-3) / 1e12))
            del rr, step_model
            torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["prod_kd", "lead", "lead_col", "teacher_large"])
    ap.add_argument("--queries", default="", help="comma-separated queries per GPU (defaults per workload)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--dropout", type=float, default=0.1, help="the reference recipes train with 0.1")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    defaults = dict(prod_kd="8,64", lead="128", lead_col="16", teacher_large="16,64")
    args.queries = [int(x) for x in (args.queries or defaults[args.workload]).split(",")]
    R = Runner()
    from simxns_b200 import _lib

    _lib.require_device()
    res = dict(prod_kd=bench_prod_kd, lead=bench_lead, lead_col=bench_lead_col, teacher_large=bench_teacher_large)[args.workload](R, args)
    if R.rank == 0 and args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        with open(args.out, "a") as f:
            for r in res:
                f.write(json.dumps(r) + "\n")
    if R.world > 1:
        R.dist.destroy_process_group()


if __name__ == "__main__":
    main()
