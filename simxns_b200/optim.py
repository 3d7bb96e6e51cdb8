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


class ShardedAdamW(AdamW):
    """AdamW with the optimizer state and the update sharded over the data-parallel ranks (ZeRO-1) for models wrapped in
    `gradsync.FlatGradSync(model, shard=True)` — the second half of SURVEY §8f N3 (the reference keeps a full replica of
    exp_avg / exp_avg_sq on every GPU and all-reduces every gradient: co_training_marco_train.py:57-69, 107-114).

    Per tower, rank r owns elements [lo, hi) of the flat parameter / gradient layout: `step()` builds the kernel table from
    the intersections of that range with the parameter tensors (so every element keeps its parameter's lr / weight decay),
    runs the same fused norm + AdamW kernels on 1/world of the elements with rank-local moments, and all-gathers the flat
    parameter buffer in place.  The gradient norm for clipping is the global one (partial sums of squares are all-reduced).
    Parameters outside the towers (a reranker head) are updated on every rank as usual.  `param_groups` (and therefore LR
    schedulers) behave as with `AdamW`; `state_dict()` holds this rank's shard of the moments only."""

    def __init__(self, sync, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.0,
                 correct_bias: bool = True, flavor: str = "hf"):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias,
                         flavor=flavor)
        if not getattr(sync, "shard", False):
            raise ValueError("ShardedAdamW needs gradsync.FlatGradSync(model, shard=True)")
        self.sync = sync
        self._step_count = 0
        self._shards = {}  # id(tower) -> (exp_avg, exp_avg_sq) of this rank's chunk
        self._tab_shard, self._tab_loose = _Table(), _Table()

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        import torch.distributed as dist

        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        mode = _lib.ADAMW_HF if self.flavor == "hf" else _lib.ADAMW_TORCH
        group_of = {}
        for group in self.param_groups:
            for p in group["params"]:
                group_of[id(p)] = group
        g0 = self.param_groups[0]
        betas, eps, correct_bias = tuple(g0["betas"]), float(g0["eps"]), bool(g0.get("correct_bias", True))
        for g in self.param_groups:
            if (tuple(g["betas"]), float(g["eps"]), bool(g.get("correct_bias", True))) != (betas, eps, correct_bias):
                raise NotImplementedError("ShardedAdamW: all groups must share betas / eps / correct_bias")
        self._step_count += 1
        rows, towers, dev = [], [], None
        for tower in self.sync._towers:
            if tower._grad_shard is None:
                continue
            lo, hi, chunk = tower._shard_range
            dev = tower._flat_param.device
            if id(tower) not in self._shards:
                self._shards[id(tower)] = (torch.zeros(chunk, dtype=torch.float32, device=dev),
                                           torch.zeros(chunk, dtype=torch.float32, device=dev))
            m, v = self._shards[id(tower)]
            for p, (off, n) in zip(tower.ordered_parameters(), tower.grad_slices()):
                a, b = max(off, lo), min(off + n, hi)
                group = group_of.get(id(p))
                if a >= b or group is None or not p.requires_grad:
                    continue
                rows.append((tower._flat_param[a:b], tower._grad_shard[a - lo:b - lo], m[a - lo:b - lo], v[a - lo:b - lo],
                             b - a, group["lr"], group["weight_decay"]))
            towers.append(tower)
        loose = []
        for p in self.sync._loose:
            group = group_of.get(id(p))
            if p.grad is None or group is None:
                continue
            _check(p, "parameter")
            _check(p.grad, "gradient")
            state = self.state[p]
            if not state:
                state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            state["step"] = self._step_count
            loose.append((p, p.grad, state["exp_avg"], state["exp_avg_sq"], p.numel(), group["lr"], group["weight_decay"]))
            dev = p.device
        if dev is None:
            return loss
        with torch.cuda.device(dev):
            _lib.require_device()
            st = torch.cuda.current_stream(dev).cuda_stream
            sq = torch.zeros((), dtype=torch.float64, device=dev)
            sq_ptr = None
            tabs = []
            if rows:
                self._tab_shard.update(rows, dev)
                tabs.append(self._tab_shard)
            if loose:
                self._tab_loose.update(loose, dev)
                tabs.append(self._tab_loose)
            if max_grad_norm is not None:
                if rows:  # this rank's chunk of the global sum of squares; the chunks partition the towers
                    _lib.check(lib.simx_grad_sq_norm(self._tab_shard.dev.data_ptr(), self._tab_shard.n, self._tab_shard.chunks,
                                                     sq.data_ptr(), st), "simx_grad_sq_norm")
                if self.sync.world > 1:
                    dist.all_reduce(sq, group=self.sync.group)
                if loose:  # replicated gradients: counted once
                    sq_l = torch.zeros((), dtype=torch.float64, device=dev)
                    _lib.check(lib.simx_grad_sq_norm(self._tab_loose.dev.data_ptr(), self._tab_loose.n, self._tab_loose.chunks,
                                                     sq_l.data_ptr(), st), "simx_grad_sq_norm")
                    sq += sq_l
                self.last_grad_norm = sq.sqrt().to(torch.float32)
                sq_ptr = sq.data_ptr()
            for tab in tabs:
                _lib.check(lib.simx_adamw_step(tab.dev.data_ptr(), tab.n, tab.chunks, mode, betas[0], betas[1], eps,
                                               self._step_count, int(correct_bias), sq_ptr, float(max_grad_norm or 0.0), 0, st),
                           "simx_adamw_step")
            touched = [p for p, *_ in loose]
            for tower in towers:  # every rank now needs everybody's updated chunk
                lo, hi, chunk = tower._shard_range
                if self.sync.world > 1:
                    dist.all_gather_into_tensor(tower._flat_param[:tower._flat_padded_numel], tower._flat_param[lo:lo + chunk],
                                                group=self.sync.group)
                tower._grad_shard = None
                touched.extend(p for p in tower.ordered_parameters() if isinstance(p, torch.nn.Parameter))
        if touched:
            torch.autograd.graph.increment_version(touched)
        return loss

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        for tower in self.sync._towers:
            tower._grad_shard = None
