"""ORACLE tooling (build container only) — runs THE REFERENCE'S OWN `train()`
(/root/reference/SimANS/co_training/co_training_marco_train.py:82-306, imported unmodified) on the CPU with the reference's
own model classes, and records what it did: the logged per-update losses, the learning rates, and every parameter after
the last update.  tests/test_gpu_driver_trace.py replays the same batches through `simxns_b200.train` on the B200 and must
land on the same numbers — the driver-level counterpart of the kernel parity tests.

What is shimmed, OUTSIDE the product package, because this image lacks it or because it is I/O:
  * `transformers.AdamW` (removed in transformers 5): a 20-line restatement of the v4 optimizer (same update as
    oracle/optim_numpy.adamw_hf and as simxns_b200.optim.AdamW(flavor="hf")), installed before the driver is imported;
  * `Rocketqa_v2Dataset` / `DataLoader` / `RandomSampler` / `tqdm`: replaced in the driver's namespace by a list of
    pre-built synthetic batches in the collate layout (the data path is pinned separately: tests/golden/batch_layout.npz);
  * `SummaryWriter`, `_save_checkpoint`, `_save_teacher_checkpoint`, `torch.distributed.barrier`: no-ops.
Nothing inside `train()` is touched: optimizer construction (`get_optimizer`), the schedulers, `set_seed`, the retriever
branch (:195-224), the accumulate / clip / step / zero_grad tail (:246-263) and the train_flag state machine run as shipped.
Dropout is set to 0 in the models (the trace must be deterministic); tiny BERT shapes keep the CPU run to seconds.

    python oracle/make_driver_trace.py        ->  tests/golden/driver_trace.npz
"""
from __future__ import annotations

import json
import logging
import math
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/SimANS")
GOLD = ROOT / "tests" / "golden"

from oracle.bert_numpy import bert_config, synth_params, synth_tokens  # noqa: E402

CFG = dict(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=400, max_pos=64)
NQ, GROUP, Q_LEN, P_LEN, CE_LEN = 3, 4, 12, 24, 32
N_BATCHES = 9


class AdamWv4(torch.optim.Optimizer):
    """transformers v4 `optimization.AdamW` (correct_bias=True): m, v EMAs; step = lr * sqrt(1-b2^t)/(1-b1^t);
    p -= step * m / (sqrt(v) + eps); then p -= lr * wd * p"""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))

    def step(self, closure=None):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p.data)
                    state["exp_avg_sq"] = torch.zeros_like(p.data)
                b1, b2 = group["betas"]
                state["step"] += 1
                state["exp_avg"].mul_(b1).add_(p.grad.data, alpha=1.0 - b1)
                state["exp_avg_sq"].mul_(b2).addcmul_(p.grad.data, p.grad.data, value=1.0 - b2)
                denom = state["exp_avg_sq"].sqrt().add_(group["eps"])
                step_size = group["lr"]
                if group["correct_bias"]:
                    step_size = step_size * math.sqrt(1.0 - b2 ** state["step"]) / (1.0 - b1 ** state["step"])
                p.data.addcdiv_(state["exp_avg"], denom, value=-step_size)
                if group["weight_decay"] > 0.0:
                    p.data.add_(p.data, alpha=-group["lr"] * group["weight_decay"])


def make_batches():
    """synthetic batches in the layout of Rocketqa_v2Dataset.get_collate_fn (MARCO_until_new.py:216-260):
    {'student': [q_ids, q_mask, p_ids, p_mask], 'teacher': [ce_ids [B,1+n,L], ce_mask, tgt]}"""
    out = []
    for i in range(N_BATCHES):
        q_ids, q_mask = synth_tokens(NQ, Q_LEN, CFG["vocab"], seed=10 + i, ragged=True)
        p_ids, p_mask = synth_tokens(NQ * GROUP, P_LEN, CFG["vocab"], seed=50 + i, ragged=True, min_len=6)
        ce_ids, ce_mask = synth_tokens(NQ * GROUP, CE_LEN, CFG["vocab"], seed=90 + i, ragged=True, min_len=8)
        tgt = np.zeros((NQ, GROUP), np.int64)
        tgt[:, 0] = 1
        out.append(dict(q_ids=q_ids, q_mask=q_mask, p_ids=p_ids, p_mask=p_mask,
                        ce_ids=ce_ids.reshape(NQ, GROUP, CE_LEN), ce_mask=ce_mask.reshape(NQ, GROUP, CE_LEN), tgt=tgt))
    return out


def main():
    sys.path.insert(0, str(REF))
    import importlib

    models = importlib.import_module("model.models")  # first: this import re-creates the lazy `transformers` module
    import transformers

    transformers.AdamW = AdamWv4
    sys.modules.setdefault("faiss", types.ModuleType("faiss"))
    drv = importlib.import_module("co_training.co_training_marco_train")

    from transformers import BertConfig

    def hf_cfg():
        return BertConfig(vocab_size=CFG["vocab"], hidden_size=CFG["hidden"], num_hidden_layers=CFG["num_layers"],
                          num_attention_heads=CFG["heads"], intermediate_size=CFG["intermediate"],
                          max_position_embeddings=CFG["max_pos"], type_vocab_size=2, hidden_dropout_prob=0.0,
                          attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12)

    def load(module, params, prefix=""):
        sd = {prefix + k: torch.from_numpy(v.copy()) for k, v in params.items()}
        missing, unexpected = module.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected

    cfg = bert_config(**CFG)
    pq, pc, pt = synth_params(cfg, seed=21), synth_params(cfg, seed=22), synth_params(cfg, seed=23)
    model = models.BiBertEncoder.__new__(models.BiBertEncoder)
    torch.nn.Module.__init__(model)
    model.question_model = models.HFBertEncoder(hf_cfg())
    model.ctx_model = models.HFBertEncoder(hf_cfg())
    load(model.question_model, pq)
    load(model.ctx_model, pc)
    teacher = models.Reranker(models.HFBertEncoder(hf_cfg()), CFG["hidden"])
    load(teacher.encoder, pt)
    head_w = (0.05 * np.random.RandomState(24).standard_normal((1, CFG["hidden"]))).astype(np.float32)
    with torch.no_grad():
        teacher.qa_classifier.weight.copy_(torch.from_numpy(head_w))
        teacher.qa_classifier.bias.zero_()

    batches = make_batches()
    t = torch.from_numpy
    collated = [dict(student=[t(b["q_ids"]), t(b["q_mask"]), t(b["p_ids"]), t(b["p_mask"])],
                     teacher=[t(b["ce_ids"]), t(b["ce_mask"]), t(b["tgt"])]) for b in batches]

    class FakeDataset(list):
        def __init__(self, *a, **k):
            super().__init__(collated)

        @staticmethod
        def get_collate_fn(args):
            return None

    drv.Rocketqa_v2Dataset = FakeDataset
    drv.RandomSampler = lambda ds: None
    drv.DataLoader = lambda ds, **k: list(ds)
    drv.tqdm = lambda it, **k: it
    drv.SummaryWriter = lambda **k: types.SimpleNamespace(add_scalar=lambda *a, **k: None, close=lambda: None)
    drv._save_checkpoint = lambda *a, **k: None
    drv._save_teacher_checkpoint = lambda *a, **k: None
    torch.distributed.barrier = lambda *a, **k: None

    logs = []

    class Grab(logging.Handler):
        def emit(self, record):
            msg = record.getMessage()
            if msg.startswith("{") and '"step"' in msg:
                logs.append(json.loads(msg))

    drv.logger.setLevel(logging.INFO)
    drv.logger.addHandler(Grab())

    args = types.SimpleNamespace(
        log_dir="/tmp/simx_trace_tb", device=torch.device("cpu"), per_gpu_train_batch_size=NQ, n_gpu=0, weight_decay=0.01,
        learning_rate=1e-3, teacher_learning_rate=5e-4, adam_epsilon=1e-8, optimizer="adamW", fp16=False, local_rank=-1, rank=0,
        world_size=1, max_steps=8, iteration_step=8, iteration_reranker_step=1, origin_data_dir="", ann_dir="", output_dir="/tmp",
        number_neg=GROUP - 1, passage_path="", gradient_accumulation_steps=2, logging_steps=1, save_steps=0, seed=42,
        scale_simmila=False, temperature_distill=2.0, max_grad_norm=0.5)
    torch.set_num_threads(8)
    final_step = drv.train(args, model, teacher, tokenizer=None, global_step=0)

    res = dict(cfg=np.array([CFG["num_layers"], CFG["hidden"], CFG["heads"], CFG["intermediate"], CFG["vocab"], CFG["max_pos"]]),
               seeds=np.array([21, 22, 23, 24]), head_w=head_w, final_step=np.array([final_step]),
               hparams=np.array([args.learning_rate, args.teacher_learning_rate, args.adam_epsilon, args.weight_decay,
                                 args.temperature_distill, args.max_grad_norm, args.gradient_accumulation_steps,
                                 args.max_steps, args.iteration_step, args.iteration_reranker_step], np.float64),
               log_step=np.array([l["step"] for l in logs]), log_loss=np.array([l["loss"] for l in logs]),
               log_distill=np.array([l["distill_loss"] for l in logs]), log_contr=np.array([l["contr_loss"] for l in logs]),
               log_lr=np.array([l["learning_rate"] for l in logs]))
    for i, b in enumerate(batches):
        for k, v in b.items():
            res[f"b{i}/{k}"] = v
    for n, p in model.named_parameters():
        res["student/" + n] = p.detach().numpy().copy()
    for n, p in teacher.named_parameters():
        res["teacher/" + n] = p.detach().numpy().copy()
    np.savez_compressed(GOLD / "driver_trace.npz", **res)
    print("reference train() ran", final_step, "updates over", len(collated), "batches; logs:")
    for l in logs:
        print("  ", l)


if __name__ == "__main__":
    main()
