"""A few fused clip + AdamW steps over two bert-base towers (for `ncu -k regex:adamw_kernel|grad_sq_norm_kernel`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from simxns_b200 import optim as sx  # noqa: E402
from simxns_b200.encoder import EncoderConfig  # noqa: E402
from simxns_b200.models import BiBertEncoder, HFBertEncoder  # noqa: E402

ec = EncoderConfig()
m = BiBertEncoder(None, question_model=HFBertEncoder(ec), ctx_model=HFBertEncoder(ec)).cuda()
for p in m.parameters():
    p.grad = torch.randn_like(p) * 1e-3
opt = sx.AdamW(m.parameters(), lr=1e-5, eps=1e-8, weight_decay=0.01, flavor="hf")
for _ in range(4):
    opt.step(max_grad_norm=1.0)
torch.cuda.synchronize()
print("ok")
