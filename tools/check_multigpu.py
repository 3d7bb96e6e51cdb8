"""Multi-GPU checks of the widened rows (run under torchrun, one rank per GPU, NCCL):
  1. sharded FlatIPIndex: every rank adds its slice of a corpus, search() == search over the whole corpus on one GPU;
  2. DDP + in-batch InfoNCE over all-gathered [CLS] rows + fused clip/AdamW: ranks hold identical parameters after 3 steps
     and the global loss equals the single-process loss on the concatenated batch (fp32-class precision).
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multigpu.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import bert_numpy as bn  # noqa: E402
from simxns_b200 import _lib  # noqa: E402
from simxns_b200.encoder import EncoderConfig  # noqa: E402
from simxns_b200.losses import fused_score_loss  # noqa: E402
from simxns_b200.models import BiBertEncoder, HFBertEncoder  # noqa: E402
from simxns_b200.optim import AdamW  # noqa: E402
from simxns_b200.search import FlatIPIndex  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()

    # ---- 1. sharded exact search ----
    g = torch.Generator().manual_seed(0)
    N, H, k = 200_003, 128, 200
    P = torch.randn(N, H, generator=g)
    Q = torch.randn(16, H, generator=g)
    cuts = [N * r // world for r in range(world + 1)]
    shard = FlatIPIndex(H, row_offset=cuts[rank])
    shard.add(P[cuts[rank]:cuts[rank + 1]])
    D, I = shard.search(Q, k)
    whole = FlatIPIndex(H)
    whole.add(P)
    Dw, Iw = whole.search_local(Q, k)
    assert torch.equal(I, Iw), "sharded ids differ from the single-index ids"
    assert torch.allclose(D, Dw, rtol=0, atol=1e-5 * float(Dw.abs().max()))
    if rank == 0:
        print(f"sharded search over {world} ranks == single index: OK (top-{k} of {N} rows, {Q.shape[0]} queries)")

    # ---- 2. DDP step with the fused optimizer ----
    ec = EncoderConfig(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=400, max_pos=64)
    torch.manual_seed(0)

    def make():
        torch.manual_seed(0)
        return BiBertEncoder(None, question_model=HFBertEncoder(ec, precision=_lib.PREC_BF16X3),
                             ctx_model=HFBertEncoder(ec, precision=_lib.PREC_BF16X3)).to(dev)

    model = make()
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=False)
    opt = AdamW(model.parameters(), lr=1e-3, eps=1e-8, weight_decay=0.01)
    nq, group = 4, 8
    t = lambda a: torch.from_numpy(a).to(dev)
    losses = []
    for step in range(3):
        batches = []
        for r in range(world):
            q_ids, q_mask = bn.synth_tokens(nq, 12, ec.vocab, seed=10 * step + r, ragged=True)
            p_ids, p_mask = bn.synth_tokens(nq * group, 24, ec.vocab, seed=100 + 10 * step + r, ragged=True, min_len=6)
            batches.append((q_ids, q_mask, p_ids, p_mask))
        q, c = ddp(*[t(a) for a in batches[rank]])
        loss, _ = fused_score_loss(q, c, ["nll_global"], group_size=group, gather=True)
        loss.backward()
        if step == 0:  # single-process loss AND gradients on the concatenated batch, same initial weights
            ref = make()
            cat = [np.concatenate([b[i] for b in batches]) for i in range(4)]
            qr, cr = ref(*[t(a) for a in cat])
            lr_, _ = fused_score_loss(qr, cr, ["nll_global"], group_size=group, gather=False)
            assert abs(lr_.item() - loss.item()) < 1e-4, (lr_.item(), loss.item())
            lr_.backward()
            # every rank differentiates the same global loss w.r.t. its own rows only, DDP averages: world * ddp grad == grad
            worst = 0.0
            gmax = max(p.grad.abs().max().item() for p in ref.parameters())
            for (n, p), (_, q) in zip(ref.named_parameters(), model.named_parameters()):
                e = (world * q.grad - p.grad).abs().max().item() / max(p.grad.abs().max().item(), 1e-3 * gmax)
                worst = max(worst, e)
            assert worst < 1e-3, f"DDP gradient differs from the single-process gradient by {worst:.2e}"
            if rank == 0:
                print(f"world x DDP-averaged gradient == single-process gradient of the global loss: worst {worst:.1e}: OK")
        opt.step(max_grad_norm=1.0)
        ddp.zero_grad()
        losses.append(loss.item())
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], x) for x in gathered), "ranks diverged after DDP + fused AdamW"
    if rank == 0:
        print(f"DDP + all-gathered InfoNCE + fused clip/AdamW on {world} ranks: parameters identical, losses {losses}: OK")

    # ---- 2b. FlatGradSync (one all-reduce per tower over its flat gradient buffer) == DDP, incl. a shared tower used twice
    #          per step, a head outside the towers, and gradient accumulation over two micro-steps ----
    from simxns_b200.gradsync import FlatGradSync
    from simxns_b200.models import Reranker

    class Both(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(3)
            self.bi = BiBertEncoder(None, question_model=HFBertEncoder(ec, precision=_lib.PREC_BF16X3))  # share_weight
            self.rr = Reranker(HFBertEncoder(ec, precision=_lib.PREC_BF16X3), ec.hidden)

        def forward(self, q, qm, p, pm):
            qe, ce = self.bi(q, qm, p, pm)
            return qe, ce, self.rr(p.view(nq, group, -1), pm.view(nq, group, -1))

    def run(wrapper):
        m = Both().to(dev)
        w = wrapper(m)
        for micro in range(2):
            q_ids, q_mask = bn.synth_tokens(nq, 12, ec.vocab, seed=700 + 10 * micro + rank, ragged=True)
            p_ids, p_mask = bn.synth_tokens(nq * group, 24, ec.vocab, seed=800 + 10 * micro + rank, ragged=True, min_len=6)
            qe, ce, logits = w(t(q_ids), t(q_mask), t(p_ids), t(p_mask))
            loss, _ = fused_score_loss(qe, ce, ["nll_global"], group_size=group, gather=True)
            loss = loss + torch.nn.functional.cross_entropy(logits, torch.zeros(nq, dtype=torch.long, device=dev))
            loss.backward()
        torch.cuda.synchronize()
        return {n: p.grad.clone() for n, p in m.named_parameters()}

    g_ddp = run(lambda m: torch.nn.parallel.DistributedDataParallel(m, device_ids=[local]))
    g_flat = run(lambda m: FlatGradSync(m))
    gmax = max(v.abs().max().item() for v in g_ddp.values())
    # two runs of the same backward differ in the order of the split-K fp32 atomics (~1e-6 of the gradient scale)
    worst = max((g_flat[n] - g_ddp[n]).abs().max().item() / max(g_ddp[n].abs().max().item(), 1e-2 * gmax) for n in g_ddp)
    assert worst < 2e-4, f"FlatGradSync gradients differ from DDP's by {worst:.2e}"
    if rank == 0:
        print(f"FlatGradSync == DistributedDataParallel (shared tower, reranker head, 2 accumulated micro-steps): worst {worst:.1e}: OK")

    # ---- 2c. sharded optimizer (ZeRO-1 on the flat layout): FlatGradSync(shard=True) + ShardedAdamW == FlatGradSync + AdamW ----
    from simxns_b200.train import get_optimizer
    import types

    oargs = types.SimpleNamespace(optimizer="adamW")

    def train3(shard):
        m = Both().to(dev)
        w = FlatGradSync(m, shard=shard)
        opt = get_optimizer(oargs, m, weight_decay=0.01, lr=1e-3, eps=1e-8, sync=w if shard else None)
        norms = []
        for step in range(3):
            for micro in range(2):  # two accumulated micro-steps per update
                q_ids, q_mask = bn.synth_tokens(nq, 12, ec.vocab, seed=900 + 20 * step + 10 * micro + rank, ragged=True)
                p_ids, p_mask = bn.synth_tokens(nq * group, 24, ec.vocab, seed=950 + 20 * step + 10 * micro + rank, ragged=True,
                                                min_len=6)
                qe, ce, logits = w(t(q_ids), t(q_mask), t(p_ids), t(p_mask))
                loss, _ = fused_score_loss(qe, ce, ["nll_global"], group_size=group, gather=True)
                loss = loss + torch.nn.functional.cross_entropy(logits, torch.zeros(nq, dtype=torch.long, device=dev))
                loss.backward()
            opt.step(max_grad_norm=0.1)  # below every norm of this run: the clip is active
            norms.append(opt.last_grad_norm.item())
            opt.zero_grad()
        torch.cuda.synchronize()
        state_bytes = sum(t_.numel() * 4 for st_ in opt.state.values() for t_ in st_.values() if torch.is_tensor(t_))
        state_bytes += sum(a.numel() * 4 + b.numel() * 4 for a, b in getattr(opt, "_shards", {}).values())
        return {n: p.detach().clone() for n, p in m.named_parameters()}, norms, state_bytes

    p_full, n_full, b_full = train3(False)
    p_shard, n_shard, b_shard = train3(True)
    assert all(abs(a - b) < 1e-4 * max(1.0, abs(a)) for a, b in zip(n_full, n_shard)), (n_full, n_shard)
    worst, worst_name = 0.0, ""
    for n in p_full:
        # Adam normalises: tensors whose true gradient is zero (key biases: softmax shift invariance; the last LayerNorm bias
        # of a tower feeding a softmax-row loss; the head bias) move by +-lr on fp32 round-off in either run — skip them and
        # compare the bulk of every other tensor
        if "key.bias" in n or n.endswith(f"layer.{ec.num_layers - 1}.output.LayerNorm.bias") or n.endswith("qa_classifier.bias") \
                or "pooler" in n:
            continue
        d = (p_full[n] - p_shard[n]).abs().mean().item() / 1e-3
        if d > worst:
            worst, worst_name = d, n
    assert worst < 2e-2, f"sharded optimizer diverges from the replicated one: mean |dp| / lr = {worst:.2e} at {worst_name}"
    flat = torch.cat([p.reshape(-1) for p in p_shard.values()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], x) for x in gathered), "ranks diverged under the sharded optimizer"
    if rank == 0:
        print(f"ShardedAdamW (reduce-scatter + 1/{world} of the Adam state: {b_shard / 2**20:.1f} vs {b_full / 2**20:.1f} MiB per rank) == "
              f"replicated AdamW after 3 clipped updates: gradient norms {['%.4f' % x for x in n_shard]}, mean |dp|/lr {worst:.1e}: OK")

    # ---- 3. the training configuration: dropout + gradient checkpointing + bf16 under DDP (ranks must stay in lock-step) ----
    torch.manual_seed(1)
    m2 = BiBertEncoder(None, question_model=HFBertEncoder(ec, precision=_lib.PREC_BF16),
                       ctx_model=HFBertEncoder(ec, precision=_lib.PREC_BF16)).to(dev)
    for tower in (m2.question_model, m2.ctx_model):
        tower.hidden_dropout_prob = tower.attention_probs_dropout_prob = 0.1
        tower.gradient_checkpointing_enable()
    ddp2 = FlatGradSync(m2)
    opt2 = AdamW(m2.parameters(), lr=1e-3, eps=1e-8, weight_decay=0.01)
    seeds = []
    for step in range(2):
        q_ids, q_mask = bn.synth_tokens(nq, 12, ec.vocab, seed=500 + 10 * step + rank, ragged=True)
        p_ids, p_mask = bn.synth_tokens(nq * group, 24, ec.vocab, seed=600 + 10 * step + rank, ragged=True, min_len=6)
        q, c = ddp2(t(q_ids), t(q_mask), t(p_ids), t(p_mask))
        seeds.append(m2.ctx_model.last_dropout_seed)
        loss, _ = fused_score_loss(q, c, ["nll_global"], group_size=group, gather=True)
        loss.backward()
        opt2.step(max_grad_norm=1.0)
        ddp2.zero_grad()
    flat = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], x) for x in gathered), "ranks diverged with dropout + checkpointing"
    all_seeds = [None] * world
    dist.all_gather_object(all_seeds, seeds)
    assert len({tuple(s) for s in all_seeds}) == world, "ranks drew the same dropout seeds"
    if rank == 0:
        print(f"dropout 0.1 + gradient checkpointing + bf16 under FlatGradSync on {world} ranks: parameters identical, "
              f"per-rank dropout seeds distinct: OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
