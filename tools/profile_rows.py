"""One encoder layer forward + backward at the passage-tower shape (for `ncu -k regex:ln_fwd_kernel|ln_bwd_kernel`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from simxns_b200 import _lib  # noqa: E402
from simxns_b200.encoder import EncoderConfig  # noqa: E402
from simxns_b200.models import HFBertEncoder  # noqa: E402

B, S = (int(sys.argv[1]) if len(sys.argv) > 1 else 2048), 128
m = HFBertEncoder(EncoderConfig(num_layers=1), precision=_lib.PREC_BF16).cuda().train()
ids = torch.randint(1000, 30000, (B, S), device="cuda")
mask = torch.ones_like(ids)
for _ in range(2):
    cls, _ = m.encode(ids, mask)
    cls[:, -1].sum().backward()
    m.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("ok")
