"""Optimizer tail of the train step on libsimx_b200: `clip_grad_norm_` and `AdamW` with the reference's call
shapes, each ONE kernel launch over every parameter tensor instead of a handful of launches per tensor.

Replaces (drop-in):
  * `torch.nn.utils.clip_grad_norm_(model.parameters(), args.max_grad_norm)`
    (/root/reference/SimANS/co_training/co_training_marco_train.py:246-259, LEAD/run_LEAD.py:232-238);
  * `AdamW(optimizer_grouped_parameters, lr=lr, eps=eps)` of `get_optimizer`
    (co_training_marco_train.py:57-69 — transformers.AdamW, `flavor="hf"`; LEAD/util.py:901 — torch.optim.AdamW,
    `flavor="torch"`).  `param_groups` / `state_dict()` / LR schedulers (`get_linear_schedule_with_warmup` writes
    `group["lr"]`) work as with the stock classes; per-parameter state keys are `step`, `exp_avg`, `exp_avg_sq`.

There is no CPU path: parameters and gradients must be fp32 CUDA tensors on a B200.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Union

import torch

from . import _lib


class _Table:
    """device copy of a `simx_opt_tensor` table; rebuilt when any pointer / lr / weight decay changed"""

    def __init__(self):
        self.key = None
        self.dev: Optional[torch.Tensor] = None
        self.n = 0
        self.chunks = 0

    def update(self, rows, device) -> None:
        """rows: list of (p, g, m, v, numel, lr, wd) with tensors or None"""
        key = tuple((_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), n, float(lr), float(wd))
                    for p, g, m, v, n, lr, wd in rows)
        if key == self.key and self.dev is not None and self.dev.device == device:
            return
        arr = (_lib.OptTensor * len(rows))()
        for i, (pp, gp, mp, vp, n, lr, wd) in enumerate(key):
            arr[i].p, arr[i].g, arr[i].m, arr[i].v = pp or None, gp or None, mp or None, vp or None
            arr[i].numel, arr[i].lr, arr[i].weight_decay = n, lr, wd
        chunks = _lib.lib().simx_opt_plan(arr, len(rows))
        if chunks <= 0:
            raise RuntimeError("simx_opt_plan failed: " + _lib.last_error())
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.dev = host.to(device)
        self.key, self.n, self.chunks = key, len(rows), chunks


def _check(p: torch.Tensor, what: str) -> None:
    if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
        raise RuntimeError(f"simxns_b200.optim: {what} must be a contiguous fp32 CUDA tensor (got {p.dtype} on {p.device}); "
                           "there is no CPU fallback")


_clip_tables = {}


def clip_grad_norm_(parameters: Union[torch.Tensor, Iterable[torch.Tensor]], max_norm: float,
                    norm_type: float = 2.0) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ (L2): returns the total gradient norm (0-dim fp32 CUDA tensor, no host sync) and
    scales every `.grad` in place by min(1, max_norm / (norm + 1e-6)).  Two launches for any number of tensors."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("only the L2 norm the reference drivers use is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return torch.zeros(())
    dev = params[0].device
    rows = []
    for p in params:
        _check(p.grad, "gradient")
        rows.append((p.grad, p.grad, None, None, p.grad.numel(), 0.0, 0.0))  # `p` slot unused by these two kernels
    tab = _clip_tables.setdefault(dev, _Table())
    with torch.cuda.device(dev):
        _lib.require_device()
        tab.update(rows, dev)
        lib = _lib.lib()
        st = torch.cuda.current_stream(dev).cuda_stream
        sq = torch.empty((), dtype=torch.float64, device=dev)
        _lib.check(lib.simx_grad_sq_norm(tab.dev.data_ptr(), tab.n, tab.chunks, sq.data_ptr(), st), "simx_grad_sq_norm")
        _lib.check(lib.simx_clip_grads(tab.dev.data_ptr(), tab.n, tab.chunks, sq.data_ptr(), float(max_norm), st),
                   "simx_clip_grads")
    return sq.sqrt().to(torch.float32)


class AdamW(torch.optim.Optimizer):
    """Fused AdamW.  `flavor="hf"`: transformers.AdamW semantics (`correct_bias` as there); `flavor="torch"`:
    torch.optim.AdamW semantics.  `step(max_grad_norm=...)` additionally folds global-norm clipping into the update (the
    norm is reduced on the device and never visits the host); `last_grad_norm` then holds it."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.0,
                 correct_bias: bool = True, flavor: str = "hf"):
        if flavor not in ("hf", "torch"):
            raise ValueError("flavor must be 'hf' (transformers.AdamW) or 'torch' (torch.optim.AdamW)")
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        self.flavor = flavor
        self._tables: List[_Table] = []
        self.last_grad_norm: Optional[torch.Tensor] = None

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        mode = _lib.ADAMW_HF if self.flavor == "hf" else _lib.ADAMW_TORCH
        # groups that share (betas, eps, correct_bias, step count) go out in one launch: bucket by those
        buckets = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p, "parameter")
                _check(p.grad, "gradient")
                if p.grad.is_sparse:
                    raise RuntimeError("AdamW does not support sparse gradients")
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                key = (p.device, tuple(group["betas"]), float(group["eps"]), bool(group.get("correct_bias", True)),
                       int(state["step"]))
                buckets.setdefault(key, []).append((p, p.grad, state["exp_avg"], state["exp_avg_sq"], p.numel(),
                                                    group["lr"], group["weight_decay"]))
        while len(self._tables) < len(buckets):
            self._tables.append(_Table())
        if max_grad_norm is not None and len(buckets) > 1:
            # parameters that differ in betas / eps / step count / device need several launches: the norm must still be
            # the global one, so clip first (one table over every gradient), then update without the fused clip
            self.last_grad_norm = clip_grad_norm_([r[0] for rows in buckets.values() for r in rows], max_grad_norm)
            max_grad_norm = None
        touched = []
        for tab, ((dev, betas, eps, correct_bias, step), rows) in zip(self._tables, buckets.items()):
            with torch.cuda.device(dev):
                _lib.require_device()
                tab.update(rows, dev)
                st = torch.cuda.current_stream(dev).cuda_stream
                sq_ptr = None
                if max_grad_norm is not None:
                    sq = torch.empty((), dtype=torch.float64, device=dev)
                    _lib.check(lib.simx_grad_sq_norm(tab.dev.data_ptr(), tab.n, tab.chunks, sq.data_ptr(), st),
                               "simx_grad_sq_norm")
                    self.last_grad_norm = sq.sqrt().to(torch.float32)
                    sq_ptr = sq.data_ptr()
                _lib.check(lib.simx_adamw_step(tab.dev.data_ptr(), tab.n, tab.chunks, mode, betas[0], betas[1], eps, step,
                                               int(correct_bias), sq_ptr, float(max_grad_norm or 0.0), 1, st),
                           "simx_adamw_step")
            touched.extend(r[0] for r in rows)
        # the kernel wrote through raw pointers: tell autograd / the encoder's bf16 weight cache (keyed on `_version`)
        if touched:
            torch.autograd.graph.increment_version(touched)
        return loss
