"""ORACLE — test / baseline infrastructure only (never imported by the product path).

The reference's train step as the reference itself runs it on a CPU: its thin wrapper classes restated over the
installed `transformers.BertModel` (the un-vendored dependency that holds the reference's arithmetic), PyTorch CPU
ops, all host threads.  /root/reference does not exist on the GPU box, so these ~40 lines are a "port"
(`cpu_baseline.kind = "port"`); in the build container tests/test_ref_torch_port.py checks them against
the real classes imported from /root/reference.

  HFBertEncoderRef  <- /root/reference/SimANS/model/models.py:58-82
  BiBertEncoderRef  <- /root/reference/SimANS/model/models.py:85-118
  nll_global_loss   <- BiEncoderNllLoss.calc, models.py:468-505 (positives at i*(1+n): LEAD/util.py:445)
  simans_kl_loss    <- /root/reference/SimANS/co_training/co_training_marco_train.py:199-214
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


def _hf_config(cfg):
    from transformers import BertConfig

    return BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["num_layers"],
                      num_attention_heads=cfg["heads"], intermediate_size=cfg["intermediate"],
                      max_position_embeddings=cfg["max_pos"], type_vocab_size=cfg["type_vocab"],
                      hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=cfg["ln_eps"])


def make_encoder(cfg):
    from transformers import BertModel

    class HFBertEncoderRef(BertModel):
        def forward(self, **kwargs):
            result = super().forward(**kwargs)
            sequence_output = result.last_hidden_state + 0 * result.pooler_output.sum()
            pooled_output = sequence_output[:, 0, :]
            return sequence_output, pooled_output, None

    return HFBertEncoderRef(_hf_config(cfg))


class BiBertEncoderRef(nn.Module):
    def __init__(self, cfg, share_weight=False):
        super().__init__()
        self.question_model = make_encoder(cfg)
        self.ctx_model = self.question_model if share_weight else make_encoder(cfg)

    def query_emb(self, input_ids, attention_mask):
        return self.question_model(input_ids=input_ids, attention_mask=attention_mask)[1]

    def body_emb(self, input_ids, attention_mask):
        return self.ctx_model(input_ids=input_ids, attention_mask=attention_mask)[1]

    def forward(self, query_ids, attention_mask_q, input_ids_a, attention_mask_a):
        return self.query_emb(query_ids, attention_mask_q), self.body_emb(input_ids_a, attention_mask_a)


def nll_global_loss(q, c, group):
    scores = torch.matmul(q, c.t())
    lsm = F.log_softmax(scores, dim=1)
    target = torch.arange(0, q.size(0) * group, group, dtype=torch.long)
    return F.nll_loss(lsm, target, reduction="mean")


def simans_kl_loss(q, c, teacher_logits, temperature=1.0, eps=1e-7):
    cc = c.reshape(q.size(0), c.size(0) // q.size(0), -1)
    sim = torch.einsum("bh,bdh->bd", q, cc)
    p_s = F.softmax(sim, dim=1)
    p_t = F.softmax(teacher_logits / temperature, dim=1)
    return torch.nn.KLDivLoss(reduction="batchmean")((p_s + eps).log(), p_t)


def load_params(module, params, prefix=""):
    sd = {prefix + k: torch.from_numpy(np.asarray(v).copy()) for k, v in params.items()}
    module.load_state_dict(sd, strict=False)


def time_cpu_step(cfg, n_queries, group, q_len, p_len, steps, warmup, threads=None, loss="nll"):
    """Times fwd + loss + bwd of the reference port on the host CPU.  Returns dict(pairs_per_s, s_per_step, cores)."""
    import os

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = BiBertEncoderRef(cfg)
    model.train()  # dropout probabilities are 0 in the config: same math as eval
    g = torch.Generator().manual_seed(1234)
    q_ids = torch.randint(1000, cfg["vocab"], (n_queries, q_len), generator=g)
    p_ids = torch.randint(1000, cfg["vocab"], (n_queries * group, p_len), generator=g)
    q_ids[:, 0] = 101
    p_ids[:, 0] = 101
    q_mask, p_mask = torch.ones_like(q_ids), torch.ones_like(p_ids)
    teacher = 3 * torch.randn(n_queries, group, generator=g)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        q, c = model(q_ids, q_mask, p_ids, p_mask)
        l = nll_global_loss(q, c, group) if loss == "nll" else simans_kl_loss(q, c, teacher)
        l.backward()
        model.zero_grad(set_to_none=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    s = float(np.mean(times))
    return dict(pairs_per_s=n_queries * group / s, s_per_step=s, cores=threads, times=times)
