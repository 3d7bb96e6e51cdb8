"""Drop-in for /root/reference/PROD/ProD_KD/model/models.py on the distillation hot path
(`from model.models import ...` in PROD/ProD_KD/run_progressive_distill_marco.py:27): the shared encoder / bi-encoder
classes, PROD's Reranker (with the `binary` head and the 3-tuple return) and its two KD loss classes with every KD_type.

PROD's `HFBertEncoder.init_encoder(args, role=None, dropout=0.1, model_type=None, number_layers=0)` picks the checkpoint AND
the depth by role (PROD/ProD_KD/model/models.py:20-72): student -> args.model_type / args.student_num_hidden_layers,
teacher -> args.teacher_model_type / args.teacher_num_hidden_layers, double_teacher -> args.double_teacher_pretrain /
args.double_teacher_num_hidden_layers; role None -> the given model_type truncated to number_layers.  The reference maps
three model names to hard-coded local directories and exits on anything else; here the name is resolved like any HF
name / local directory.
"""
from typing import Optional

from torch import nn

from . import models as _m
from .losses import BiEncoderKDLoss, BiEncoderNllLoss, CrossBERTKDLoss  # noqa: F401
from .models import dot_product_scores, init_weights  # noqa: F401
from .models import ProdReranker as Reranker  # noqa: F401

__all__ = ["HFBertEncoder", "BiBertEncoder", "Reranker", "BiEncoderNllLoss", "CrossBERTKDLoss", "BiEncoderKDLoss",
           "dot_product_scores", "init_weights"]

_ROLE_FIELDS = {"student": ("model_type", "student_num_hidden_layers"),
                "teacher": ("teacher_model_type", "teacher_num_hidden_layers"),
                "double_teacher": ("double_teacher_pretrain", "double_teacher_num_hidden_layers")}


class HFBertEncoder(_m.HFBertEncoder):
    @classmethod
    def init_encoder(cls, args, role: Optional[str] = None, dropout: float = 0.1, model_type=None, number_layers: int = 0):
        if role is not None and role not in _ROLE_FIELDS:
            raise ValueError(f"no such type role: {role!r}")  # the reference prints this and exits
        if model_type is None:
            if role is None:
                raise ValueError("no such type role: init_encoder needs a role or a model_type")
            model_type = getattr(args, _ROLE_FIELDS[role][0])
        cfg = _m._load_hf_config(model_type, "bert")
        if dropout != 0:
            cfg.attention_probs_dropout_prob = dropout
            cfg.hidden_dropout_prob = dropout
        cfg.gradient_checkpointing = getattr(args, "gradient_checkpointing", False)
        layers = number_layers if role is None else getattr(args, _ROLE_FIELDS[role][1])
        if not layers or layers <= 0:
            raise ValueError(f"num_hidden_layers for role {role!r} must be positive, got {layers!r}")
        cfg.num_hidden_layers = int(layers)
        return cls.from_pretrained(model_type, config=cfg)


class BiBertEncoder(_m.BiBertEncoder):
    """PROD/ProD_KD/model/models.py:208-256: `BiBertEncoder(args, role)` builds both towers for that role"""

    def __init__(self, args, role: Optional[str] = None, question_model: Optional[nn.Module] = None,
                 ctx_model: Optional[nn.Module] = None):
        nn.Module.__init__(self)
        if question_model is not None:
            self.question_model = question_model
            self.ctx_model = ctx_model if ctx_model is not None else question_model
            return
        share = hasattr(args, 'share_weight') and args.share_weight
        if role is not None:
            self.role = role
            if role == 'student' and args.model_type == 'distilbert-base-uncased':
                make = lambda: _m.HFDistilBertEncoder.init_encoder(args, pretrained_model_name=args.model_type)
            else:
                make = lambda: HFBertEncoder.init_encoder(args, role)
        else:  # the reference's default when no role is given (models.py:230-235)
            make = lambda: HFBertEncoder.init_encoder(args, model_type='nghuyong/ernie-2.0-base-en', number_layers=12)
        self.question_model = make()
        self.ctx_model = self.question_model if share else make()
