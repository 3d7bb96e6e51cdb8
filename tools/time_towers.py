"""Standalone cost of each tower of the headline step (one B200): query tower [128 x 32], passage tower [2048 x 128],
forward + backward with a random [CLS] cotangent, CUDA events after warm-up."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from simxns_b200 import _lib  # noqa: E402
from simxns_b200.encoder import EncoderConfig  # noqa: E402
from simxns_b200.models import HFBertEncoder  # noqa: E402


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


res = {}
for name, B, S in (("query", 128, 32), ("passage", 2048, 128)):
    m = HFBertEncoder(EncoderConfig(), precision=_lib.PREC_BF16).cuda().train()
    ids = torch.randint(1000, 30000, (B, S), device="cuda")
    mask = torch.ones_like(ids)
    d = torch.randn(B, 768, device="cuda")

    def step():
        cls, _ = m.encode(ids, mask)
        (cls[:, -1] * d).sum().backward()
        m.zero_grad(set_to_none=True)

    def fwd():
        with torch.no_grad():
            m.encode(ids, mask)

    res[name + "_fwd_bwd_ms"] = timed(step)
    res[name + "_fwd_ms"] = timed(fwd)
    del m
print(json.dumps({k: round(v, 3) for k, v in res.items()}))
