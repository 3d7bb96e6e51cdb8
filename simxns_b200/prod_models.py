"""Drop-in for /root/reference/PROD/ProD_KD/model/models.py on the distillation hot path
(`from model.models import ...` in PROD/ProD_KD/run_progressive_distill_marco.py:27): the shared encoder / bi-encoder
classes, PROD's Reranker (with the `binary` head and the 3-tuple return) and its two KD loss classes with every KD_type."""
from .losses import BiEncoderKDLoss, BiEncoderNllLoss, CrossBERTKDLoss  # noqa: F401
from .models import BiBertEncoder, HFBertEncoder, dot_product_scores, init_weights  # noqa: F401
from .models import ProdReranker as Reranker  # noqa: F401

__all__ = ["HFBertEncoder", "BiBertEncoder", "Reranker", "BiEncoderNllLoss", "CrossBERTKDLoss", "BiEncoderKDLoss",
           "dot_product_scores", "init_weights"]
