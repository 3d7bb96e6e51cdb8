"""Data-parallel gradient exchange for models whose towers run on libsimx_b200 — the role of the reference's
`torch.nn.parallel.DistributedDataParallel` wrap (/root/reference/SimANS/co_training/co_training_marco_train.py:107-114),
laid out for how these towers produce gradients.

Every backward pass of a tower writes ALL of its gradients into one flat fp32 buffer (simx_encoder_grad_offset).  Stock DDP
copies the 199 tensors into ~35 buckets of 25 MB and all-reduces them one by one; here each tower's buffer is reduced IN
PLACE by a single NCCL all-reduce (ReduceOp.AVG) issued the moment that tower's pass completes — the passage tower's
exchange runs under the query tower's backward — and `p.grad` is then pointed at (or accumulated from) views of the reduced
buffer: no bucket copies, one large message per tower instead of many small ones.  Parameters outside the towers (the
Reranker head, RobertaDot's projection) travel in one extra flat all-reduce at the end of backward.

    model = FlatGradSync(BiBertEncoder(args).to(device))      # instead of DistributedDataParallel(model, ...)
    q, c = model(...); loss.backward(); optimizer.step()

Same contract as DDP for the loop body: gradients are averaged over ranks and ready when `backward()` returns (stream
order), `model.module` is the wrapped module, parameters are broadcast from rank 0 at construction, gradient accumulation
over micro-steps adds.  `compress="bf16"` halves the bytes on the wire (gradients are rounded to bf16 for the exchange
only — a deviation from the reference's fp32 all-reduce, off by default).

`shard=True` (second half) turns the exchange into ZeRO-1 over the same flat layout: each tower's parameters
are moved into ONE flat fp32 buffer laid out like its gradient buffer (`p.data` become views: names, shapes and
`state_dict()` are unchanged), the gradient buffer is REDUCE-SCATTERED in place (rank r ends up with the averaged chunk r
only), `optim.ShardedAdamW` updates chunk r of the parameters with rank-local Adam moments (1/world of the optimizer state
and of the update work) and all-gathers the flat parameter buffer in place.  Tower parameters then carry no `.grad`.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn

from .encoder import B200EncoderBase


class FlatGradSync(nn.Module):
    def __init__(self, module: nn.Module, process_group=None, compress: Optional[str] = None, broadcast: bool = True,
                 shard: bool = False):
        super().__init__()
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("FlatGradSync needs an initialised torch.distributed process group")
        if compress not in (None, "bf16"):
            raise ValueError("compress must be None or 'bf16'")
        if shard and compress:
            raise ValueError("shard=True exchanges fp32 chunks; compress is not combined with it")
        self.module = module
        self.group = process_group
        self.compress = compress
        self.shard = bool(shard)
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self._towers: List[B200EncoderBase] = []
        seen = set()
        tower_params = set()
        for m in module.modules():
            if isinstance(m, B200EncoderBase) and id(m) not in seen:
                seen.add(id(m))
                self._towers.append(m)
                m._grad_sink = self._sink
                tower_params.update(id(p) for p in m.parameters())
        self._loose = [p for p in module.parameters() if id(p) not in tower_params]
        self._pending = []
        self._callback_queued = False
        if broadcast and self.world > 1:
            with torch.no_grad():
                tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
                for t in tensors:  # same initial weights everywhere, as DDP's constructor guarantees
                    dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
            for m in self._towers:
                m.invalidate_weights()
        if self.shard:
            for m in self._towers:
                self._flatten_tower(m)

    # ---- shard=True: parameters of a tower live in one flat buffer with the layout of its gradient buffer -----------
    def _flatten_tower(self, tower: B200EncoderBase) -> None:
        slices = tower.grad_slices()
        numel = max(off + n for off, n in slices)
        chunk = (-(-numel // self.world) + 63) // 64 * 64          # per-rank chunk, a multiple of 64 elements
        padded = chunk * self.world
        params = tower.ordered_parameters()
        flat = torch.zeros(padded, dtype=torch.float32, device=params[0].device)
        with torch.no_grad():
            for p, (off, n) in zip(params, slices):
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
        tower._flat_param = flat
        tower._flat_padded_numel = padded
        tower._shard_range = (self.rank * chunk, min((self.rank + 1) * chunk, numel), chunk)  # [lo, hi), chunk size
        tower._grad_shard = None
        tower.invalidate_weights()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ---- called by encoder._TowerPass.finish() from inside backward ---------------------------------------------------
    def _sink(self, tower: B200EncoderBase, st) -> None:
        flat = st.flat
        work, wire = None, None
        if self.shard:
            lo, _, chunk = tower._shard_range
            padded = tower._flat_padded_numel
            if self.world > 1:  # in place: rank r's chunk of the buffer receives the average of everybody's chunk r
                work = dist.reduce_scatter_tensor(flat[lo:lo + chunk], flat[:padded], op=dist.ReduceOp.AVG, group=self.group,
                                                  async_op=True)
        elif self.world > 1:
            if self.compress == "bf16":
                wire = flat.to(torch.bfloat16)
                work = dist.all_reduce(wire, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            else:
                work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        self._pending.append((tower, st.params, flat, wire, work))
        if not self._callback_queued:
            self._callback_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

    # ---- end of backward ----------------------------------------------------------------------------------------------
    def _finalize(self) -> None:
        self._callback_queued = False
        pending, self._pending = self._pending, []
        for tower, params, flat, wire, work in pending:
            if work is not None:
                work.wait()  # stream-level: the current stream waits for the collective, the host does not
            if wire is not None:
                flat.copy_(wire)
            if self.shard:  # keep this rank's chunk only (accumulated over micro-steps); tower parameters get no .grad
                lo, _, chunk = tower._shard_range
                mine = flat[lo:lo + chunk]
                if tower._grad_shard is None:
                    if getattr(tower, "_grad_shard_buf", None) is None:
                        tower._grad_shard_buf = torch.empty_like(mine)  # persistent: the optimizer's kernel table keeps its address
                    tower._grad_shard = tower._grad_shard_buf.copy_(mine)
                else:
                    tower._grad_shard.add_(mine)
                continue
            for p, (off, n) in zip(params, tower.grad_slices()):
                if not isinstance(p, nn.Parameter) or not p.requires_grad:
                    continue
                g = flat[off:off + n].view(p.shape)
                if p.grad is None:
                    p.grad = g
                else:
                    p.grad.add_(g)  # gradient accumulation / shared towers (two passes per step)
        loose = [p for p in self._loose if p.requires_grad and p.grad is not None]
        if loose and self.world > 1:
            buf = torch.cat([p.grad.reshape(-1) for p in loose])
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
            off = 0
            for p in loose:
                p.grad.copy_(buf[off:off + p.numel()].view_as(p.grad))
                off += p.numel()

    def detach_sinks(self) -> None:
        """towers hand their gradients to autograd again (no exchange) until attach_sinks()"""
        for m in self._towers:
            m._grad_sink = None

    def attach_sinks(self) -> None:
        for m in self._towers:
            m._grad_sink = self._sink
