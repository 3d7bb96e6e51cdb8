"""Exact inner-product search on libsimx_b200 (`simx_topk_*`) behind the faiss calls of the SimANS ANN refresh
(/root/reference/SimANS/co_training/co_training_generate.py:359-384 `get_new_faiss_index`, :412-427 `get_question_topk`):

    cpu_index = faiss.IndexFlatIP(dim); gpu_index = faiss.index_cpu_to_all_gpus(cpu_index, co)   # co.shard = True
    gpu_index.add(passage_embedding.astype(np.float32)); D, I = gpu_index.search(question_embedding, 200)

becomes `index = FlatIPIndex(dim); index.add(emb); D, I = index.search(q, 200)`.  One process per GPU: each rank `add`s
its own shard of the corpus (rows `[row_offset, row_offset + n)` of the global numbering, the role of co.shard) and
`search` merges the per-rank top-k with one all-gather of [nq, k] (scores, ids) followed by the same select kernel.
The shard stays resident in HBM as three bf16 planes (fp32-class operands of the tcgen05 score GEMM).  No CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class FlatIPIndex:
    def __init__(self, d: int, device: Optional[torch.device] = None, capacity: int = 0, row_offset: int = 0):
        if d <= 0 or d % 64:
            raise ValueError(f"dimension {d} must be a positive multiple of 64")
        self.d = d
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FlatIPIndex lives on a B200: there is no CPU fallback")
        self.ntotal = 0
        self.row_offset = int(row_offset)  # global id of this shard's row 0
        self._cap = 0
        self._planes: Optional[torch.Tensor] = None
        self._pending = []  # fp32 slabs added before the shard size is known (capacity == 0)
        if capacity:
            self._allocate(capacity)

    # ---- building -------------------------------------------------------------------------------------------
    def _allocate(self, cap: int) -> None:
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.require_device()
            self._planes = torch.empty(lib.simx_topk_corpus_bytes(cap, self.d), dtype=torch.uint8, device=self.device)
        self._cap = cap

    def _to_dev(self, x) -> torch.Tensor:
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"expected [n, {self.d}] embeddings, got {tuple(x.shape)}")
        return x

    def add(self, x) -> None:
        """faiss `index.add(x)`: append rows.  With a preset `capacity` rows are converted in place slab by slab;
        otherwise they are staged and laid out at the first `search` (the plane stride depends on the final size)."""
        x = self._to_dev(x)
        if self._cap:
            if self.ntotal + x.shape[0] > self._cap:
                raise ValueError(f"index capacity {self._cap} exceeded")
            self._write(x, self.ntotal, self._cap)
        else:
            self._pending.append(x)
        self.ntotal += x.shape[0]

    def _write(self, x: torch.Tensor, at: int, total: int) -> None:
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().simx_topk_add(x.data_ptr(), x.shape[0], at, total, self.d, self._planes.data_ptr(),
                                                torch.cuda.current_stream(self.device).cuda_stream), "simx_topk_add")

    def _finalize(self) -> None:
        if self._pending:
            self._allocate(self.ntotal)
            at = 0
            for x in self._pending:
                self._write(x, at, self._cap)
                at += x.shape[0]
            self._pending = []
        elif self._cap and self.ntotal != self._cap:
            raise RuntimeError(f"index allocated for {self._cap} rows holds {self.ntotal}: add the rest before searching")

    # ---- searching ------------------------------------------------------------------------------------------
    def search_local(self, q, k: int, max_queries: int = 4096) -> Tuple[torch.Tensor, torch.Tensor]:
        """top-k of this rank's shard: (D [nq, k] fp32 descending, I [nq, k] int64 GLOBAL row ids, -1 padding)"""
        if self.ntotal == 0:
            raise RuntimeError("search on an empty index")
        self._finalize()
        q = self._to_dev(q)
        lib = _lib.lib()
        nq = q.shape[0]
        D = torch.empty(nq, k, dtype=torch.float32, device=self.device)
        I = torch.empty(nq, k, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            step = min(max_queries, nq)
            ws = torch.empty(lib.simx_topk_workspace_bytes(step, self.ntotal, self.d, k), dtype=torch.uint8, device=self.device)
            for b in range(0, nq, step):
                n = min(step, nq - b)
                _lib.check(lib.simx_topk_search(q[b:b + n].data_ptr(), n, self._planes.data_ptr(), self.ntotal, self.d, k,
                                                D[b:b + n].data_ptr(), I[b:b + n].data_ptr(), ws.data_ptr(), st),
                           "simx_topk_search")
        if self.row_offset:
            I = torch.where(I >= 0, I + self.row_offset, I)
        return D, I

    def search(self, q, k: int, group=None, as_numpy: bool = False):
        """faiss `index.search(q, k)` over ALL shards: every rank passes the same queries and gets the same (D, I)."""
        D, I = self.search_local(q, k)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            D, I = merge_shard_results(D, I, k, group)
        if as_numpy:
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I


def merge_shard_results(D: torch.Tensor, I: torch.Tensor, k: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """all-gather the per-rank [nq, k] results and keep the best k per query, (score desc, id asc).  The gathered
    candidates are only world*k per query, so the merge is a sort of a tiny matrix."""
    world = dist.get_world_size(group)
    nq = D.shape[0]
    Dg = D.new_empty((world * nq, k))  # rank-major concatenation along dim 0 (the layout every backend accepts)
    Ig = I.new_empty((world * nq, k))
    dist.all_gather_into_tensor(Dg, D.contiguous(), group=group)
    dist.all_gather_into_tensor(Ig, I.contiguous(), group=group)
    Dc = Dg.view(world, nq, k).permute(1, 0, 2).reshape(nq, world * k)
    Ic = Ig.view(world, nq, k).permute(1, 0, 2).reshape(nq, world * k)
    return merge_candidates(Dc, Ic, k)


def merge_candidates(Dc: torch.Tensor, Ic: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """best k of [nq, m] candidates by (score desc, id asc); padding ids (-1) sort last"""
    ids = torch.where(Ic >= 0, Ic, torch.full_like(Ic, torch.iinfo(torch.int64).max))
    order = torch.argsort(ids, dim=1, stable=True)
    Dc, Ic = torch.gather(Dc, 1, order), torch.gather(Ic, 1, order)
    order = torch.argsort(Dc, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(Dc, 1, order), torch.gather(Ic, 1, order)


@torch.no_grad()
def embed_passages(encode_fn, input_ids: torch.Tensor, batch_size: int = 2048) -> torch.Tensor:
    """`embed_passages` of the reference (co_training_generate.py:97-121) without the tokenizer: encode already
    tokenised rows [n, S] batch by batch with `encode_fn(input_ids=..., attention_mask=...)` (`model.body_emb` /
    `model.query_emb`) under no_grad — the inference path of the encoder (nothing saved) — and return [n, H] on the GPU
    (the reference moves every batch to the CPU and concatenates there; the index here is fed straight from HBM).
    A sequence's embedding does not depend on its batch (bit-exact, tests/test_gpu_parity.py), so the reference's
    batch of 256 is only a scheduling choice: 2048 rows per call fill the 148 SMs' tile waves (35k -> 45k passages/s)."""
    out = []
    for b in range(0, input_ids.shape[0], batch_size):
        ids = input_ids[b:b + batch_size]
        out.append(encode_fn(input_ids=ids, attention_mask=(ids != 0).long()))
    return torch.cat(out, dim=0)
