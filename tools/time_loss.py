"""Time the global InfoNCE (fused cooperative launch vs the unfused 3-GEMM checker path) at the headline per-GPU size
[128, 2048] and the 8-GPU size [1024, 16384]:  python tools/time_loss.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from simxns_b200.losses import fused_score_loss  # noqa: E402

for GQ, GC, nq in ((128, 2048, 128), (1024, 16384, 128)):
    q = (torch.randn(GQ, 768, device="cuda") * 0.1).requires_grad_(True)
    c = (torch.randn(GC, 768, device="cuda") * 0.1).requires_grad_(True)
    for unf in (False, True):
        f = lambda: fused_score_loss(q, c, ["nll_global"], group_size=GC // GQ, gather=False, unfused_check=unf)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        print(f"[{GQ}, {GC}] x 768 (all rows / columns local)  {'unfused (3 GEMMs + row kernel)' if unf else 'fused (1 launch)'}: "
              f"{e0.elapsed_time(e1) / 20:.3f} ms")
