"""SimANS ambiguous-negative sampling on the GPU (libsimx_b200 `simx_simans_sample`).

Reference: `Rocketqa_v2Dataset.__getitem__` (/root/reference/SimANS/utils/MARCO_until_new.py:165-202) and
`TraditionDataset` (/root/reference/SimANS/utils/util_wiki.py:613-639) run this per query on the CPU inside DataLoader
workers; here a whole batch of queries is sampled in one launch from a device-resident score table."""
from __future__ import annotations

import torch

from . import _lib


def simans_sample(neg_scores: torch.Tensor, pos_score: torch.Tensor, k: int, *, tau: float = None, a: float = None,
                  b: float = 0.0, n_valid: torch.Tensor = None, seed: int = 0) -> torch.Tensor:
    """neg_scores [NQ, NC] (retriever scores of the candidates, NC <= 1024), pos_score [NQ] -> indices [NQ, k] (int32).
    Give `tau` for the MS-MARCO form exp(-tau |s - s+|) (tau = 3 in the reference) or `a`, `b` for the NQ/TQ form
    exp(-a (s - s+ + b)^2)."""
    if (tau is None) == (a is None):
        raise ValueError("give either tau (mode 0) or a[, b] (mode 1)")
    if neg_scores.device.type != "cuda":
        raise RuntimeError("simans_sample needs CUDA tensors (no CPU fallback)")
    lib = _lib.lib()
    with torch.cuda.device(neg_scores.device):
        _lib.require_device()
        s = neg_scores.to(torch.float32).contiguous()
        p = pos_score.to(device=s.device, dtype=torch.float32).contiguous()
        nv = None if n_valid is None else n_valid.to(device=s.device, dtype=torch.int32).contiguous()
        NQ, NC = s.shape
        out = torch.empty(NQ, k, dtype=torch.int32, device=s.device)
        mode, p0, p1 = (0, float(tau), 0.0) if tau is not None else (1, float(a), float(b))
        _lib.check(lib.simx_simans_sample(s.data_ptr(), _lib.ptr(nv), p.data_ptr(), NQ, NC, k, mode, p0, p1,
                                          int(seed) & (2**64 - 1), out.data_ptr(),
                                          torch.cuda.current_stream(s.device).cuda_stream), "simx_simans_sample")
    return out
