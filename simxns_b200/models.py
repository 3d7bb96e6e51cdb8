"""Drop-in replacement for the reference's `model/models.py` on the dual-encoder hot path.

Same class names, constructor arguments, attributes, call signatures, return values and `state_dict` layout as
/root/reference/SimANS/model/models.py (HFBertEncoder :58-82, BiBertEncoder :85-160, Reranker :638-659,
BiEncoderNllLoss :468-514) and /root/reference/PROD/ProD_KD/model/models.py (HFBertEncoder :13-82 with
`role=` layer truncation, BiBertEncoder :208-298, CrossBERTKDLoss :668-781) — but every encoder forward/backward runs
on the sm_100a kernels of libsimx_b200.  A driver switches with one import line:

    from simxns_b200.models import BiBertEncoder, HFBertEncoder, Reranker, BiEncoderNllLoss

There is no PyTorch/CPU fallback: parameters must live on a B200 before forward() is called.
"""
from __future__ import annotations

import os

import dataclasses
import json
from pathlib import Path
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor as T
from torch import nn

from .encoder import B200EncoderBase, EncoderConfig
from .losses import BiEncoderNllLoss, CrossBERTKDLoss, simans_distill_loss  # noqa: F401  (re-exported: reference names)

__all__ = ["HFBertEncoder", "HFDistilBertEncoder", "HFRobertaEncoder", "BiBertEncoder", "BiRobertaEncoder", "RobertaDot",
           "Reranker", "ProdReranker", "BiEncoderNllLoss",
           "CrossBERTKDLoss",
           "simans_distill_loss", "dot_product_scores", "init_weights"]


# ---------------------------------------------------------------------------------------------------------
# configuration / checkpoint helpers (stand in for transformers' from_pretrained on local files)
# ---------------------------------------------------------------------------------------------------------
def _load_hf_config(model_type, arch="bert"):
    """BertConfig.from_pretrained(model_type) as the reference does (SimANS/model/models.py:69); works for local
    directories and anything in the HF cache.  Without `transformers` only a local config.json is read."""
    try:
        import transformers

        cls = {"distilbert": transformers.DistilBertConfig, "roberta": transformers.RobertaConfig}.get(
            arch, transformers.BertConfig)
        return cls.from_pretrained(model_type)
    except ImportError:
        with open(Path(model_type) / "config.json") as f:
            return SimpleNamespace(**json.load(f))


def _find_weights(model_type) -> Optional[dict]:
    """state dict of a local HF checkpoint directory / HF-cache entry (pytorch_model.bin or model.safetensors)"""
    cands = []
    p = Path(str(model_type))
    if p.is_dir():
        cands = [p / "model.safetensors", p / "pytorch_model.bin"]
    else:
        try:
            from transformers.utils import cached_file

            for name in ("model.safetensors", "pytorch_model.bin"):
                try:
                    f = cached_file(str(model_type), name)
                    if f:
                        cands.append(Path(f))
                except Exception:
                    pass
        except ImportError:
            pass
    for f in cands:
        if f.exists():
            if f.suffix == ".safetensors":
                from safetensors.torch import load_file

                return load_file(str(f))
            return torch.load(str(f), map_location="cpu", weights_only=True)
    return None


def _strip_prefix(sd: dict, own_keys) -> dict:
    """HF checkpoints store the encoder under `bert.` / `distilbert.` (task heads) or at top level"""
    own = set(own_keys)
    out = {}
    for k, v in sd.items():
        for pre in ("", "bert.", "distilbert.", "encoder.bert.", "model."):
            if k.startswith(pre) and k[len(pre):] in own:
                out[k[len(pre):]] = v
                break
    return out


def init_weights(modules):
    """SimANS/model/models.py:457-466"""
    for module in modules:
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=0.02)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()


def dot_product_scores(q_vectors, ctx_vectors):
    """SimANS/model/models.py:564-572"""
    return torch.matmul(q_vectors, torch.transpose(ctx_vectors, 0, 1))


# ---------------------------------------------------------------------------------------------------------
class HFBertEncoder(B200EncoderBase):
    """BERT tower.  `HFBertEncoder(config)` builds a randomly initialised model from an HF BertConfig (or an
    EncoderConfig); `HFBertEncoder.init_encoder(args, dropout, model_type)` mirrors the reference classmethod.

    forward(input_ids=, attention_mask=) -> (sequence_output [B,S,H], pooled_output [B,H], None) where
    pooled_output = sequence_output[:, 0, :] (the reference multiplies BertPooler's output by 0 — its parameters
    exist, are saved in state_dict and receive zero gradients: SimANS/model/models.py:77-82)."""

    def __init__(self, config, precision: Optional[int] = None):
        enc_cfg = config if isinstance(config, EncoderConfig) else EncoderConfig.from_hf(config)
        assert enc_cfg.hidden > 0, "Encoder hidden_size can't be zero"
        super().__init__(enc_cfg, precision)
        if isinstance(config, EncoderConfig):
            config = SimpleNamespace(hidden_size=enc_cfg.hidden, num_hidden_layers=enc_cfg.num_layers,
                                     num_attention_heads=enc_cfg.heads, intermediate_size=enc_cfg.intermediate,
                                     vocab_size=enc_cfg.vocab, max_position_embeddings=enc_cfg.max_pos,
                                     type_vocab_size=enc_cfg.type_vocab, layer_norm_eps=enc_cfg.ln_eps,
                                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        self.config = config
        try:
            import transformers

            self.version = int(transformers.__version__.split('.')[0])
        except ImportError:
            self.version = 4
        # training-time configuration of the reference (SimANS/model/models.py:66-74): nn.Dropout(p) at the HF sites while
        # the module is in train() mode, and --gradient_checkpointing.  Both run inside the kernels (csrc/dropout.cuh,
        # SIMX_SAVE_LAYER_INPUTS); model.eval() — the teacher under no_grad, evaluation, the parity fixtures — disables dropout.
        self.hidden_dropout_prob = float(getattr(config, "hidden_dropout_prob", 0.0) or 0.0)
        self.attention_probs_dropout_prob = float(getattr(config, "attention_probs_dropout_prob", 0.0) or 0.0)
        self.gradient_checkpointing = bool(getattr(config, "gradient_checkpointing", False))

    # -- reference constructor path -----------------------------------------------------------------------
    @classmethod
    def init_encoder(cls, args, dropout: float = 0.1, model_type=None, role: Optional[str] = None):
        if model_type is None:
            model_type = getattr(args, "model_type", None) or getattr(args, "pretrained_model_name", None)
        cfg = _load_hf_config(model_type, "bert")
        if dropout != 0:
            cfg.attention_probs_dropout_prob = dropout
            cfg.hidden_dropout_prob = dropout
        cfg.gradient_checkpointing = getattr(args, "gradient_checkpointing", False)
        # PROD: layer-truncated student / teacher (PROD/ProD_KD/model/models.py:22-40)
        if role == "student" and getattr(args, "student_num_hidden_layers", None):
            cfg.num_hidden_layers = args.student_num_hidden_layers
        if role == "teacher" and getattr(args, "teacher_num_hidden_layers", None):
            cfg.num_hidden_layers = args.teacher_num_hidden_layers
        return cls.from_pretrained(model_type, config=cfg)

    @classmethod
    def from_pretrained(cls, model_type, config=None, **_):
        if config is None:
            config = _load_hf_config(model_type, "bert")
        model = cls(config)
        sd = _find_weights(model_type)
        if sd is None:
            raise FileNotFoundError(f"no pytorch_model.bin / model.safetensors found for {model_type!r}")
        own = model.state_dict().keys()
        missing, unexpected = model.load_state_dict(_strip_prefix(sd, own), strict=False)
        if any("pooler" not in m for m in missing):
            raise RuntimeError(f"checkpoint {model_type!r} lacks encoder weights: {missing[:5]} ...")
        return model

    def init_weights(self):
        from .encoder import _hf_init_

        _hf_init_(self, self.enc_cfg.init_range)

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def gradient_checkpointing_enable(self, *a, **k):
        """HF PreTrainedModel API (CAPSTONE/models/modules.py:54): keep layer inputs only, recompute inside backward"""
        self.gradient_checkpointing = True

    def gradient_checkpointing_disable(self):
        self.gradient_checkpointing = False

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # Checkpoints written with older transformers persist two buffers (`embeddings.position_ids`,
        # `embeddings.token_type_ids`) that this module does not have.  nn.Module.load_state_dict on a PARENT
        # (BiBertEncoder, Reranker, RobertaDot — the way every reference driver loads, several with strict=True:
        # wiki/co_training_wiki_train.py:364, PROD run_progressive_distill_*.py) recurses through this hook with the
        # parent's prefix and never calls a child's load_state_dict, so the keys are dropped here.
        for k in (prefix + "embeddings.position_ids", prefix + "embeddings.token_type_ids"):
            state_dict.pop(k, None)
        return super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                             error_msgs)

    # -- reference forward ----------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, inputs_embeds=None,
                output_attentions=None, output_hidden_states=None, _want_sequence: bool = True, **kwargs):
        if token_type_ids is not None and bool((token_type_ids != 0).any()):
            raise NotImplementedError("non-zero token_type_ids: the reference never passes segment ids "
                                      "(SimANS/model/models.py:654 'TODO: provide segment values')")
        if output_attentions:
            raise NotImplementedError("attention maps are never materialised by the fused attention kernel")
        cls, seq = self.encode(input_ids, attention_mask, want_sequence=_want_sequence, inputs_embeds=inputs_embeds)
        pooled_output = cls[:, -1, :]
        self.last_cls_layers = cls  # [B, L+1, H]; per-layer [CLS] rows (LEAD-style consumers)
        return seq, pooled_output, None


class HFDistilBertEncoder(HFBertEncoder):
    """6-layer DistilBERT student of LEAD (LEAD/models.py:31-77, LEAD/modeling_distilbert.py)."""

    def __init__(self, config, precision: Optional[int] = None):
        enc_cfg = config if isinstance(config, EncoderConfig) else EncoderConfig.from_hf(config)
        if enc_cfg.arch != "distilbert":
            raise ValueError("HFDistilBertEncoder needs a DistilBertConfig")
        B200EncoderBase.__init__(self, enc_cfg, precision)
        self.config = config if not isinstance(config, EncoderConfig) else SimpleNamespace(
            hidden_size=enc_cfg.hidden, dim=enc_cfg.hidden, n_layers=enc_cfg.num_layers)
        if not hasattr(self.config, "hidden_size"):
            self.config.hidden_size = enc_cfg.hidden
        self.version = 4
        # DistilBertModel reads `dropout` / `attention_dropout` (LEAD/modeling_distilbert.py:104, :143, :240); LEAD's
        # init_encoder sets the BERT-named fields, which DistilBERT ignores (LEAD/models.py:45-47) — so the values in force
        # are those of the checkpoint's config.json (0.1 / 0.1 for distilbert-base-uncased)
        self.hidden_dropout_prob = float(getattr(config, "dropout", 0.0) or 0.0)
        self.attention_probs_dropout_prob = float(getattr(config, "attention_dropout", 0.0) or 0.0)
        self.gradient_checkpointing = bool(getattr(config, "gradient_checkpointing", False))

    @classmethod
    def init_encoder(cls, args, dropout: float = 0.1, pretrained_model_name=None):
        name = pretrained_model_name or getattr(args, "pretrained_model_name", None)
        cfg = _load_hf_config(name, "distilbert")
        cfg.gradient_checkpointing = getattr(args, "gradient_checkpointing", False)
        model = cls(cfg)
        sd = _find_weights(name)
        if sd is None:
            raise FileNotFoundError(f"no weights found for {name!r}")
        model.load_state_dict(_strip_prefix(sd, model.state_dict().keys()), strict=False)
        return model


_SIDE_STREAMS = {}


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    """one extra stream per device, shared by every model of the process (kept out of the modules: they are deep-copied)"""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx)
        # parameters live on the default stream, the query tower's gradients are produced on the side stream: intended
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
    return _SIDE_STREAMS[idx]


class BiBertEncoder(nn.Module):
    """Bi-Encoder model component. Encapsulates query/question and context/passage encoders.
    (SimANS/model/models.py:85-118; PROD passes role='student'/'teacher', PROD/ProD_KD/model/models.py:208-256)"""

    def __init__(self, args, role: Optional[str] = None, question_model: Optional[nn.Module] = None,
                 ctx_model: Optional[nn.Module] = None):
        super().__init__()
        if question_model is not None:  # explicit towers (tests, random-init benchmarks)
            self.question_model = question_model
            self.ctx_model = ctx_model if ctx_model is not None else question_model
            return
        self.question_model = HFBertEncoder.init_encoder(args, role=role)
        if hasattr(args, 'share_weight') and args.share_weight:
            self.ctx_model = self.question_model
        else:
            self.ctx_model = HFBertEncoder.init_encoder(args, role=role)

    def query_emb(self, input_ids, attention_mask):
        _, pooled_output, _ = self.question_model(input_ids=input_ids, attention_mask=attention_mask,
                                                  _want_sequence=False)
        return pooled_output

    def body_emb(self, input_ids, attention_mask):
        _, pooled_output, _ = self.ctx_model(input_ids=input_ids, attention_mask=attention_mask, _want_sequence=False)
        return pooled_output

    # The two towers are independent until the loss.  The query tower is ~1.5 % of the step's FLOPs but 3.3 % of its time when
    # run by itself (128 x 32 tokens: a third of a wave per GEMM, ~10 us kernels, 5.1 ms forward + backward): it is issued on
    # a second CUDA stream so that its kernels fill the SMs the passage tower's persistent kernels leave idle at their tails.
    # autograd runs each tower's backward on the stream its forward ran on and joins the streams when backward() returns, so
    # callers see ordinary single-stream semantics.  Off: SIMX_OVERLAP_TOWERS=0 or `model.overlap_towers = False`; never
    # used when the towers share weights (one module, one operand cache).
    overlap_towers = os.environ.get("SIMX_OVERLAP_TOWERS", "1") != "0"

    def _towers(self, query_ids, attention_mask_q, input_ids_a, attention_mask_a):
        if (self.overlap_towers and self.question_model is not self.ctx_model and query_ids.is_cuda
                and input_ids_a.device == query_ids.device):
            dev = query_ids.device
            main = torch.cuda.current_stream(dev)
            side = _side_stream(dev)
            side.wait_stream(main)  # the batch, and any optimizer step on the master weights, are complete
            with torch.cuda.stream(side):
                q_embs = self.query_emb(query_ids, attention_mask_q)
            a_embs = self.body_emb(input_ids_a, attention_mask_a)
            main.wait_stream(side)
            q_embs.record_stream(main)
            return q_embs, a_embs
        return self.query_emb(query_ids, attention_mask_q), self.body_emb(input_ids_a, attention_mask_a)

    def forward(self, query_ids, attention_mask_q, input_ids_a=None, attention_mask_a=None, input_ids_b=None,
                attention_mask_b=None):
        if input_ids_b is None:
            return self._towers(query_ids, attention_mask_q, input_ids_a, attention_mask_a)
        q_embs = self.query_emb(query_ids, attention_mask_q)
        a_embs = self.body_emb(input_ids_a, attention_mask_a)
        b_embs = self.body_emb(input_ids_b, attention_mask_b)
        logit_matrix = torch.cat([(q_embs * a_embs).sum(-1).unsqueeze(1), (q_embs * b_embs).sum(-1).unsqueeze(1)],
                                 dim=1)  # [B, 2]
        lsm = F.log_softmax(logit_matrix, dim=1)
        loss = -1.0 * lsm[:, 0]
        return (loss.mean(),)

    # the adversarial-embedding variants (SimANS/model/models.py:120-160): the hard negatives arrive as input EMBEDDINGS
    # (`inputs_embeds=`), differentiable, through the same kernels (simx_encoder_forward_embeds)
    def _adv_embs(self, query_ids, attention_mask_q, input_ids_a, attention_mask_a, input_ids_embed_b, attention_mask_b):
        q_embs = self.query_emb(query_ids, attention_mask_q)
        a_embs = self.body_emb(input_ids_a, attention_mask_a)
        _, b_embs, _ = self.ctx_model(inputs_embeds=input_ids_embed_b, attention_mask=attention_mask_b, _want_sequence=False)
        return q_embs, a_embs, b_embs

    def forward_adv_triplet(self, query_ids, attention_mask_q, input_ids_a=None, attention_mask_a=None,
                            input_ids_embed_b=None, attention_mask_b=None):
        q_embs, a_embs, b_embs = self._adv_embs(query_ids, attention_mask_q, input_ids_a, attention_mask_a,
                                                input_ids_embed_b, attention_mask_b)
        logit_matrix = torch.cat([(q_embs * a_embs).sum(-1).unsqueeze(1), (q_embs * b_embs).sum(-1).unsqueeze(1)], dim=1)
        lsm = F.log_softmax(logit_matrix, dim=1)
        loss = -1.0 * lsm[:, 0]
        return (loss.mean(),)

    def _adv_pair_logits(self, *args):
        q_embs, a_embs, b_embs = self._adv_embs(*args)
        question_num = q_embs.size(0)
        neg_local_ctx_vectors = b_embs.reshape(question_num, b_embs.size(0) // question_num, -1)
        neg_simila = torch.einsum("bh,bdh->bd", [q_embs, neg_local_ctx_vectors])
        pos_simil = (q_embs * a_embs).sum(-1).unsqueeze(1)
        return pos_simil, neg_simila

    def forward_adv_pairloss(self, query_ids, attention_mask_q, input_ids_a=None, attention_mask_a=None,
                             input_ids_embed_b=None, attention_mask_b=None):
        pos_simil, neg_simila = self._adv_pair_logits(query_ids, attention_mask_q, input_ids_a, attention_mask_a,
                                                      input_ids_embed_b, attention_mask_b)
        lsm = F.log_softmax(torch.cat([pos_simil, neg_simila], dim=1), dim=1)
        loss = -1.0 * lsm[:, 1]  # column 1, as in the reference (:144)
        return (loss.mean(), lsm)

    def forward_adv_pairloss_mse(self, query_ids, attention_mask_q, input_ids_a=None, attention_mask_a=None,
                                 input_ids_embed_b=None, attention_mask_b=None):
        pos_simil, neg_simila = self._adv_pair_logits(query_ids, attention_mask_q, input_ids_a, attention_mask_a,
                                                      input_ids_embed_b, attention_mask_b)
        loss = nn.MSELoss()(neg_simila, pos_simil)
        return (loss.mean(), 0)


class Reranker(nn.Module):
    """cross-encoder teacher: encoder + Linear(H,1) on [CLS] (SimANS/model/models.py:638-659)"""

    def __init__(self, encoder: nn.Module, hidden_size):
        super().__init__()
        self.encoder = encoder
        self.qa_classifier = nn.Linear(hidden_size, 1)

    def forward(self, input_ids: T, attention_mask: T):
        # N - number of questions in a batch, M - number of passages per question, L - sequence length
        N, M, L = input_ids.size()
        relevance_logits = self._forward(input_ids.view(N * M, L), attention_mask.view(N * M, L))
        return relevance_logits.view(N, M)

    def _forward(self, input_ids, attention_mask):
        _, pooled, _ = self.encoder(input_ids=input_ids, attention_mask=attention_mask, _want_sequence=False)
        return self.qa_classifier(pooled)  # == qa_classifier(sequence_output[:, 0, :])


class ProdReranker(Reranker):
    """PROD's cross-encoder teacher (PROD/ProD_KD/model/models.py:1232-1254): the SimANS Reranker plus a 2-way `binary`
    head on the same [CLS] row; forward returns (binary_logits [N, M, 2], relevance_logits [N, M], None).  Exported as
    `Reranker` by simxns_b200.prod_models (the drop-in for PROD's `model/models.py`)."""

    def __init__(self, encoder: nn.Module, hidden_size):
        nn.Module.__init__(self)
        self.encoder = encoder
        self.binary = nn.Linear(hidden_size, 2)
        self.qa_classifier = nn.Linear(hidden_size, 1)
        init_weights([self.binary, self.qa_classifier])

    def forward(self, input_ids: T, attention_mask: T):
        N, M, L = input_ids.size()
        binary_logits, relevance_logits, _ = self._forward(input_ids.view(N * M, L), attention_mask.view(N * M, L))
        return binary_logits.view(N, M, 2), relevance_logits.view(N, M), None

    def _forward(self, input_ids, attention_mask):
        _, pooled, _ = self.encoder(input_ids=input_ids, attention_mask=attention_mask, _want_sequence=False)
        return self.binary(pooled), self.qa_classifier(pooled), None


# ---------------------------------------------------------------------------------------------------------
# SimANS MS-Doc towers (RoBERTa): SimANS/model/models.py:168-232 (HFRobertaEncoder, BiRobertaEncoder), :262-352
# (EmbeddingMixin, BaseModelDot, RobertaDot)
# ---------------------------------------------------------------------------------------------------------
class HFRobertaEncoder(HFBertEncoder):
    """HF RobertaModel on the B200 kernels: the BERT block with one token type, LayerNorm eps 1e-5 and RoBERTa's position
    ids (pad + running count of real tokens; pad id 1 on both embedding tables).  Same forward contract as
    HFBertEncoder: (sequence_output, sequence_output[:, 0], None); `add_pooling_layer=False` drops the (dead) pooler
    parameters from the module, as RobertaDot's backbone does."""

    def __init__(self, config, precision: Optional[int] = None, add_pooling_layer: bool = True):
        if isinstance(config, EncoderConfig):
            enc_cfg = config
        else:
            enc_cfg = EncoderConfig.from_hf(config)
            if not enc_cfg.roberta_positions:  # a plain namespace without model_type: this class means RoBERTa
                enc_cfg.roberta_positions, enc_cfg.pad_token_id = True, getattr(config, "pad_token_id", 1)
        enc_cfg = dataclasses.replace(enc_cfg, pooler=bool(add_pooling_layer))
        super().__init__(enc_cfg, precision)
        if not isinstance(config, EncoderConfig):
            self.config = config
            self.hidden_dropout_prob = float(getattr(config, "hidden_dropout_prob", 0.0) or 0.0)
            self.attention_probs_dropout_prob = float(getattr(config, "attention_probs_dropout_prob", 0.0) or 0.0)
            self.gradient_checkpointing = bool(getattr(config, "gradient_checkpointing", False))

    @classmethod
    def init_encoder(cls, args, dropout: float = 0.1, model_type=None):
        if model_type is None:
            model_type = args.model_type
        cfg = _load_hf_config(model_type, "roberta")
        if dropout != 0:
            cfg.attention_probs_dropout_prob = dropout
            cfg.hidden_dropout_prob = dropout
        cfg.gradient_checkpointing = getattr(args, "gradient_checkpointing", False)
        return cls.from_pretrained(model_type, config=cfg)

    @classmethod
    def from_pretrained(cls, model_type, config=None, **kw):
        if config is None:
            config = _load_hf_config(model_type, "roberta")
        model = cls(config, **kw)
        sd = _find_weights(model_type)
        if sd is None:
            raise FileNotFoundError(f"no pytorch_model.bin / model.safetensors found for {model_type!r}")
        own = model.state_dict().keys()
        stripped = {}
        for k, v in sd.items():
            for pre in ("", "roberta.", "model.roberta."):
                if k.startswith(pre) and k[len(pre):] in own:
                    stripped[k[len(pre):]] = v
                    break
        missing, _ = model.load_state_dict(stripped, strict=False)
        if any("pooler" not in m for m in missing):
            raise RuntimeError(f"checkpoint {model_type!r} lacks encoder weights: {missing[:5]} ...")
        return model


class BiRobertaEncoder(BiBertEncoder):
    """SimANS/model/models.py:194-232: two HFRobertaEncoder towers behind the BiBertEncoder interface"""

    def __init__(self, args, question_model: Optional[nn.Module] = None, ctx_model: Optional[nn.Module] = None):
        nn.Module.__init__(self)
        if question_model is not None:
            self.question_model = question_model
            self.ctx_model = ctx_model if ctx_model is not None else question_model
            return
        self.question_model = HFRobertaEncoder.init_encoder(args)
        if hasattr(args, 'share_weight') and args.share_weight:
            self.ctx_model = self.question_model
        else:
            self.ctx_model = HFRobertaEncoder.init_encoder(args)


class RobertaDot(nn.Module):
    """SimANS/model/models.py:333-352 (+ BaseModelDot :316-331, EmbeddingMixin :262-313): one shared RoBERTa tower
    (`self.roberta`, no pooler), `embeddingHead = Linear(H, output_embedding_size)` and `norm = LayerNorm` on the first
    token (or the masked mean when `model_argobj.use_mean`).  state_dict keys: roberta.*, embeddingHead.*, norm.*"""

    def __init__(self, config, model_argobj=None, precision: Optional[int] = None):
        super().__init__()
        self.use_mean = False if model_argobj is None else bool(model_argobj.use_mean)
        self.roberta = HFRobertaEncoder(config, precision=precision, add_pooling_layer=False)
        hidden = self.roberta.enc_cfg.hidden
        self.output_embedding_size = getattr(config, "output_embedding_size", hidden)
        self.embeddingHead = nn.Linear(hidden, self.output_embedding_size)
        self.norm = nn.LayerNorm(self.output_embedding_size)
        for m in (self.embeddingHead,):  # EmbeddingMixin._init_weights: N(0, 0.02) weights (biases keep torch's default)
            m.weight.data.normal_(mean=0.0, std=0.02)

    def masked_mean(self, t, mask):
        s = torch.sum(t * mask.unsqueeze(-1).float(), axis=1)
        d = mask.sum(axis=1, keepdim=True).float()
        return s / d

    def query_emb(self, input_ids, attention_mask):
        seq, first, _ = self.roberta(input_ids=input_ids, attention_mask=attention_mask, _want_sequence=self.use_mean)
        full_emb = self.masked_mean(seq, attention_mask) if self.use_mean else first
        return self.norm(self.embeddingHead(full_emb))

    def body_emb(self, input_ids, attention_mask):
        return self.query_emb(input_ids, attention_mask)

    def forward(self, input_ids, attention_mask, is_query, *args):
        assert len(args) == 0
        return self.query_emb(input_ids, attention_mask) if is_query else self.body_emb(input_ids, attention_mask)
