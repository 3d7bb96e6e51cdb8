"""Drop-in for /root/reference/LEAD/models.py on the dual-encoder path (HFBertEncoder :144-196, HFDistilBertEncoder
:31-77, BiBertEncoder :198-242, Reranker :249-267): same class names, constructor / forward signatures and 5-tuple
returns, running on libsimx_b200.

Difference that matters: LEAD returns (and all-gathers, util.py:572-579) the FULL per-layer hidden states
`all_layer_hidden [B, L, S, H]` although for the dual-encoder and cross-encoder teachers `distill_loss` only reads the
[CLS] row `[:, l, 0, :]` of the selected layers (util.py:605, 631).  Here `all_layer_hidden` of HFBertEncoder /
HFDistilBertEncoder is `[B, L, 1, H]` by default — the [CLS] rows only — so every `[:, l, 0, :]` read in LEAD's code keeps
working and carries gradient back into layer l (the encoder backward accepts a cotangent on every layer's [CLS] row).
Set `encoder.full_layer_hidden = True` for the reference's full `[B, L, S, H]` (the query-document attention maps of
util.py:504-529 read every token row); HFColBertEncoder always returns it, because `colbert_score` on a layer
(util.py:657-660) is a late interaction over all tokens.  The full states carry gradients too (simx_encoder_hidden_state /
simx_encoder_add_cotangent).  Layer indexing is LEAD's: entry k is the output of transformer layer k (the embedding output
is not included, LEAD/modeling_bert.py:607-646).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor as T
from torch import nn

from . import models as _m
from .encoder import EncoderConfig


class HFBertEncoder(_m.HFBertEncoder):
    """LEAD/models.py:144-196"""

    @classmethod
    def init_encoder(cls, args, dropout: float = 0.1, pretrained_model_name=None):
        name = args.pretrained_model_name
        path = "pretrain_models/MASTER-MARCO" if name == "master" else name
        cfg = _m._load_hf_config(path, "bert")
        cfg.output_hidden_states = True
        if dropout != 0:
            cfg.attention_probs_dropout_prob = dropout
            cfg.hidden_dropout_prob = dropout
        cfg.gradient_checkpointing = getattr(args, "gradient_checkpointing", False)
        return cls.from_pretrained(path, config=cfg)

    init_encoder_from_my_model = init_encoder
    full_layer_hidden = False   # True: all_layer_hidden is the reference's [B, L, S, H] instead of the [CLS] rows

    def forward(self, input_ids=None, attention_mask=None, output_attentions=None, **kwargs):
        if output_attentions:
            raise NotImplementedError("attention maps (LEAD VIRT losses) are not produced by the fused attention kernel")
        if self.full_layer_hidden:
            cls, seq, hidden = self.encode(input_ids, attention_mask, want_sequence=kwargs.pop("_want_sequence", False),
                                           want_hidden=True)
            return seq, cls[:, -1, :], None, None, hidden[:, 1:]
        cls, seq = self.encode(input_ids, attention_mask, want_sequence=kwargs.pop("_want_sequence", False))
        pooled_output = cls[:, -1, :]
        all_layer_hidden = cls[:, 1:, None, :]  # [B, L, 1, H]: the [CLS] row of every transformer layer
        return seq, pooled_output, None, None, all_layer_hidden


class HFColBertEncoder(_m.HFBertEncoder):
    """LEAD/models.py:79-142: the ColBERT teacher tower — BERT, optionally `linear = nn.Linear(hidden, 768)` on the last
    hidden state (`config.add_linear`), token embeddings L2-normalised over the feature axis.
    forward(mode, device, input_ids=, attention_mask=) -> (normalize(col_out) [B,S,768], all_layer_hidden [B, L(+1), S, H],
    None): every layer's full hidden state (plus the projected output as one more "layer" with add_linear), all
    differentiable.  The reference builds a punctuation skip-list from a downloaded tokenizer in __init__ but its forward
    never applies it (`mask = None`, models.py:134); `dim` is likewise unused there."""

    def __init__(self, config, precision: Optional[int] = None, mask_punctuation: bool = False, dim: int = 128):
        super().__init__(config, precision)
        self.dim = dim
        self.add_linear = bool(getattr(config, "add_linear", False))
        if self.add_linear:
            self.linear = nn.Linear(self.enc_cfg.hidden, 768)
            _m.init_weights([self.linear])

    @classmethod
    def init_encoder(cls, args, dropout: float = 0.1, pretrained_model_name=None):
        path = args.pretrained_model_name
        cfg = _m._load_hf_config(path, "bert")
        cfg.add_linear = getattr(args, "add_linear", False)
        cfg.output_hidden_states = True
        if dropout != 0:
            cfg.attention_probs_dropout_prob = dropout
            cfg.hidden_dropout_prob = dropout
        cfg.gradient_checkpointing = getattr(args, "gradient_checkpointing", False)
        return cls.from_pretrained(path, config=cfg)

    init_encoder_from_my_model = init_encoder

    def forward(self, mode=None, device=None, input_ids=None, attention_mask=None, **kwargs):
        _, seq, hidden = self.encode(input_ids, attention_mask, want_sequence=True, want_hidden=True)
        all_layer_hidden = hidden[:, 1:]
        col_out = seq
        if self.add_linear:
            col_out = self.linear(seq)
            if self.enc_cfg.hidden != 768:
                raise NotImplementedError("add_linear with hidden != 768: the appended layer has another width than the stack")
            all_layer_hidden = torch.cat([all_layer_hidden, col_out[:, None]], dim=1)
        return torch.nn.functional.normalize(col_out, p=2, dim=2), all_layer_hidden, None


class HFDistilBertEncoder(_m.HFDistilBertEncoder):
    """LEAD/models.py:31-77.  `config.add_linear` (--add_linear): an extra `linear = nn.Linear(hidden, 768)` on the last
    hidden state; its output becomes the sequence / pooled output and is appended to the per-layer stack as one more
    "layer" (models.py:64-70) — here applied to the [CLS] rows, the only rows LEAD's losses read."""

    full_layer_hidden = False   # True: all_layer_hidden is the reference's [B, L, S, H] instead of the [CLS] rows

    def __init__(self, config, precision: Optional[int] = None):
        super().__init__(config, precision)
        self.add_linear = bool(getattr(config, "add_linear", False))
        if self.add_linear:
            self.linear = nn.Linear(self.enc_cfg.hidden, 768)
            _m.init_weights([self.linear])

    def forward(self, input_ids=None, attention_mask=None, output_attentions=None, **kwargs):
        if output_attentions:
            raise NotImplementedError("attention maps (LEAD VIRT losses) are not produced by the fused attention kernel")
        want_seq = kwargs.pop("_want_sequence", False)
        if self.add_linear and self.enc_cfg.hidden != 768:
            raise NotImplementedError("add_linear with hidden != 768: the appended layer has another width than the stack")
        if self.full_layer_hidden:  # the reference's [B, L(+1), S, H]
            cls, seq, hidden = self.encode(input_ids, attention_mask, want_sequence=True, want_hidden=True)
            all_layer_hidden = hidden[:, 1:]
            if self.add_linear:
                seq = self.linear(seq)
                all_layer_hidden = torch.cat([all_layer_hidden, seq[:, None]], dim=1)
            return (seq if want_seq or self.add_linear else None), seq[:, 0, :], None, None, all_layer_hidden
        cls, seq = self.encode(input_ids, attention_mask, want_sequence=want_seq)
        pooled_output = cls[:, -1, :]
        all_layer_hidden = cls[:, 1:, None, :]
        if self.add_linear:
            if seq is not None:
                seq = self.linear(seq)
            pooled_output = self.linear(pooled_output)
            all_layer_hidden = torch.cat([all_layer_hidden, pooled_output[:, None, None, :]], dim=1)  # [B, L+1, 1, 768]
        return seq, pooled_output, None, None, all_layer_hidden


class BiBertEncoder(nn.Module):
    """LEAD/models.py:198-242: forward(query_ids, attention_mask_q, doc_ids, attention_mask_d) ->
    (q_embs, d_embs, q_all_layer_hidden, d_all_layer_hidden, mask_doc)"""

    def __init__(self, args, question_model: Optional[nn.Module] = None, ctx_model: Optional[nn.Module] = None):
        super().__init__()
        if question_model is not None:
            self.model_type = getattr(args, "model_type", "dual_encoder") if args is not None else "dual_encoder"
            self.question_model = question_model
            self.ctx_model = ctx_model if ctx_model is not None else question_model
            return
        self.pretrained_model_name = args.pretrained_model_name
        self.model_type = args.model_type
        self.device = args.device
        enc = (HFColBertEncoder if 'colbert' in args.model_type else
               HFDistilBertEncoder if 'distilbert' in args.model_type else HFBertEncoder)
        self.question_model = enc.init_encoder(args)
        if hasattr(args, 'share_weight') and args.share_weight:
            self.ctx_model = self.question_model
        else:
            self.ctx_model = enc.init_encoder(args)

    def query_emb(self, mode, input_ids, attention_mask):
        if 'colbert' in mode:
            col_output, all_layer_hidden, _ = self.question_model(mode, getattr(self, "device", None), input_ids=input_ids,
                                                                  attention_mask=attention_mask)
            return col_output, all_layer_hidden, None
        _, pooled_output, _, _, all_layer_hidden = self.question_model(input_ids=input_ids, attention_mask=attention_mask)
        return pooled_output, all_layer_hidden, None

    def body_emb(self, mode, input_ids, attention_mask):
        if 'colbert' in mode:
            col_output, all_layer_hidden, mask_doc = self.ctx_model(mode, getattr(self, "device", None), input_ids=input_ids,
                                                                    attention_mask=attention_mask)
            return col_output, all_layer_hidden, mask_doc
        _, pooled_output, _, _, all_layer_hidden = self.ctx_model(input_ids=input_ids, attention_mask=attention_mask)
        return pooled_output, all_layer_hidden, None

    def forward(self, query_ids, attention_mask_q, doc_ids=None, attention_mask_d=None):
        q_embs, q_all_layer_hidden, _ = self.query_emb(self.model_type + '_query', query_ids, attention_mask_q)
        d_embs, d_all_layer_hidden, mask_doc = self.body_emb(self.model_type + '_doc', doc_ids, attention_mask_d)
        return q_embs, d_embs, q_all_layer_hidden, d_all_layer_hidden, mask_doc


class Reranker(nn.Module):
    """LEAD/models.py:249-267: also returns the per-layer relevance logits `[L, N, M]` (qa_classifier on every layer's
    [CLS]); attention maps are returned as None."""

    def __init__(self, encoder: nn.Module, hidden_size):
        super().__init__()
        self.encoder = encoder
        self.qa_classifier = nn.Linear(hidden_size, 1)
        _m.init_weights([self.qa_classifier])

    def forward(self, input_ids: T, attention_mask: T):
        N, M, L = input_ids.size()
        seq, pooled, _, _, all_layer_hidden = self.encoder(input_ids=input_ids.view(N * M, L),
                                                           attention_mask=attention_mask.view(N * M, L))
        rank_logits = self.qa_classifier(pooled)
        rank_logits_all_layer = self.qa_classifier(all_layer_hidden[:, :, 0, :])  # [N*M, L, 1]
        return rank_logits.view(N, M), None, rank_logits_all_layer.permute([1, 0, 2]).view(rank_logits_all_layer.shape[1], N, M)
