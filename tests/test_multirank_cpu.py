"""world_size-2 gloo test (CPU) of the multi-GPU host logic: the [CLS]-embedding all-gather layout (rank-major), the
per-rank row/column ownership handed to the loss kernel, and the fact that the per-rank gradients of the sharded step
add up to the single-process gradient (remote slices are constants, no second collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import losses_numpy as ln


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simxns_b200.losses import gather_embeddings

    nq, group, H = 3, 8, 16
    rng = np.random.RandomState(100 + rank)
    q_local = torch.from_numpy(rng.standard_normal((nq, H)).astype(np.float32))
    c_local = torch.from_numpy(rng.standard_normal((nq * group, H)).astype(np.float32))
    Q = gather_embeddings(q_local)
    Cm = gather_embeddings(c_local)
    assert Q.shape == (world * nq, H) and Cm.shape == (world * nq * group, H)
    assert torch.equal(Q[rank * nq:(rank + 1) * nq], q_local)          # rank-major layout
    assert torch.equal(Cm[rank * nq * group:(rank + 1) * nq * group], c_local)
    # what _ScoreLossFn passes to the kernel for this rank
    out, dQ, dC = ln.score_loss(Q.numpy(), Cm.numpy(), terms=["nll_global"], q_begin=rank * nq, q_count=nq,
                                c_begin=rank * nq * group, group=group, pos_idx=np.arange(world * nq) * group)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), Q=Q.numpy(), C=Cm.numpy(), dQ=dQ, dC=dC, loss=out["total"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_gradient_ownership(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{i}.npz") for i in range(world)]
    assert np.array_equal(r[0]["Q"], r[1]["Q"]) and np.array_equal(r[0]["C"], r[1]["C"])
    assert abs(r[0]["loss"] - r[1]["loss"]) < 1e-12  # every rank evaluates the same global loss
    full, dQ, dC = ln.score_loss(r[0]["Q"], r[0]["C"], terms=["nll_global"], group=8,
                                 pos_idx=np.arange(r[0]["Q"].shape[0]) * 8)
    assert abs(full["total"] - r[0]["loss"]) < 1e-12
    assert np.allclose(np.concatenate([r[0]["dQ"], r[1]["dQ"]]), dQ, atol=1e-12)
    assert np.allclose(np.concatenate([r[0]["dC"], r[1]["dC"]]), dC, atol=1e-12)


def _search_worker(rank, world, port, out_dir):
    """sharded index merge (faiss co.shard=True): each rank holds rows [cut_r, cut_{r+1}) and its local top-k (from the
    oracle here — the kernel needs a GPU); merge_shard_results must reproduce the global top-k on every rank"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import search_numpy as sn
    from simxns_b200.search import merge_shard_results

    rng = np.random.RandomState(7)
    # integer-valued embeddings: every inner product is exact in float64 whatever the BLAS blocking, so ties are real ties
    P, Q = rng.randint(-8, 9, (500, 32)).astype(np.float64), rng.randint(-8, 9, (5, 32)).astype(np.float64)
    P[300:320] = P[10:30]  # exact ties across the shard boundary: the lower global row must win
    cuts = [0, 260, 500]
    k = 40
    D, I, _ = sn.flat_ip_search(P[cuts[rank]:cuts[rank + 1]], Q, k)
    I = I + cuts[rank]
    Dm, Im = merge_shard_results(torch.from_numpy(D), torch.from_numpy(I), k)
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), D=Dm.numpy(), I=Im.numpy(), P=P, Q=Q)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_merge(tmp_path):
    from oracle import search_numpy as sn

    world = 2
    mp.spawn(_search_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"s{i}.npz") for i in range(world)]
    D, I, _ = sn.flat_ip_search(r[0]["P"], r[0]["Q"], 40)
    for x in r:
        assert np.array_equal(x["I"], I)        # bit-exact ids (ties -> lowest global row) on every rank
        assert np.allclose(x["D"], D, atol=1e-12)


def _posidx_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simxns_b200.losses import global_positive_indices

    n_local_ctx = 12
    local = [0, 5, 9] if rank == 0 else [2, 4, 11]  # positives need not sit at i*(1+n) (run_de_model.py:203-220)
    g = global_positive_indices(local, n_local_ctx)
    np.save(os.path.join(out_dir, f"p{rank}.npy"), g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_positive_index_bookkeeping(tmp_path):
    """the reference's loop: positive_idx_per_question.extend([v + total_ctxs for v in positive_idx]); total_ctxs += n"""
    world = 2
    mp.spawn(_posidx_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want, total = [], 0
    for positive_idx in ([0, 5, 9], [2, 4, 11]):
        want.extend([v + total for v in positive_idx])
        total += 12
    for r in range(world):
        assert np.load(tmp_path / f"p{r}.npy").tolist() == want


def _flat_worker(rank, world, port, out_dir):
    """gradsync.FlatGradSync host logic without a GPU: the constructor's broadcast and the shard=True parameter layout"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.gradsync import FlatGradSync
    from simxns_b200.models import BiBertEncoder, HFBertEncoder

    torch.manual_seed(100 + rank)  # different initial weights per rank: the constructor must make them equal
    cfg = EncoderConfig(vocab=97, hidden=64, num_layers=2, heads=1, intermediate=128, max_pos=40, type_vocab=2)
    model = BiBertEncoder(None, question_model=HFBertEncoder(cfg), ctx_model=HFBertEncoder(cfg))
    before = {k: v.clone() for k, v in model.state_dict().items()}
    sync = FlatGradSync(model, shard=True)
    after = model.state_dict()
    assert list(after.keys()) == list(before.keys())            # names and order survive the flattening
    out = {"keys": np.array(len(after))}
    for i, tower in enumerate((model.question_model, model.ctx_model)):
        flat, padded = tower._flat_param, tower._flat_padded_numel
        lo, hi, chunk = tower._shard_range
        slices = tower.grad_slices()
        numel = max(off + n for off, n in slices)
        assert padded == chunk * world and chunk % 64 == 0 and padded >= numel
        assert lo == rank * chunk and hi == min((rank + 1) * chunk, numel)
        for p, (off, n) in zip(tower.ordered_parameters(), slices):
            assert p.data_ptr() == flat[off:off + n].data_ptr() and p.numel() == n    # parameters ARE views of the buffer
        covered = np.zeros(numel, dtype=np.int32)
        for off, n in slices:
            covered[off:off + n] += 1
        assert covered.max() == 1                                # slices never overlap
        out[f"flat{i}"] = flat.numpy().copy()
        out[f"range{i}"] = np.array([lo, hi, chunk, numel])
    out["sum"] = np.array(sum(float(v.double().sum()) for v in after.values()))
    np.savez(os.path.join(out_dir, f"f{rank}.npz"), **out)
    assert sync.module is model
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradsync_layout(tmp_path):
    world = 2
    mp.spawn(_flat_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"f{i}.npz") for i in range(world)]
    for i in range(2):
        assert np.array_equal(r[0][f"flat{i}"], r[1][f"flat{i}"])   # rank 0's weights everywhere after construction
        (lo0, hi0, chunk, numel), (lo1, hi1, _, _) = r[0][f"range{i}"], r[1][f"range{i}"]
        assert lo0 == 0 and hi0 == lo1 and hi1 == numel              # the two chunks tile [0, numel)
    assert r[0]["sum"] == r[1]["sum"]


def _col_worker(rank, world, port, out_dir):
    """losses.gather_tensor + lead_col_distill_loss(gather=True) across ranks (pure tensor code: runs on gloo / CPU)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simxns_b200.losses import gather_tensor, lead_col_distill_loss

    g = np.load(os.path.join(out_dir, "inputs.npz"))
    nq, group = 2, 3                                       # per rank
    sl_q, sl_c = slice(rank * nq, (rank + 1) * nq), slice(rank * nq * group, (rank + 1) * nq * group)
    t = {k: torch.from_numpy(g[k][sl_q if k.startswith("q") else sl_c].copy()).requires_grad_(True)
         for k in ("q_db", "c_db", "q_col", "c_col", "qL_db", "cL_db", "qL_col", "cL_col")}
    mask = torch.from_numpy(g["doc_mask"][sl_c].copy())
    full = gather_tensor(t["q_db"])
    assert full.shape[0] == world * nq and torch.equal(full[sl_q], t["q_db"])
    loss, _ = lead_col_distill_loss(t["q_db"], t["c_db"], t["q_col"], t["c_col"], mask, t["qL_db"], t["cL_db"], t["qL_col"],
                                    t["cL_col"], selected_db=[0, 2], selected_col=[1, 3], temperature=2.0,
                                    layer_temperature=0.5, para_dis=1.3, reweight=True, train_col=True, gather=True)
    loss.backward()
    np.savez(os.path.join(out_dir, f"c{rank}.npz"), loss=np.array([loss.item()]), **{k: v.grad.numpy() for k, v in t.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_col_distill_equals_single_process(tmp_path):
    """every rank evaluates the SAME global loss; its local gradients are the matching slices of the single-process
    gradients (remote rows are constants: LEAD/util.py:200-204)"""
    from simxns_b200.losses import lead_col_distill_loss

    rng = np.random.RandomState(5)
    GQ, group, H, Sq, Sd = 4, 3, 16, 4, 6
    mk = lambda *s: (rng.standard_normal(s) * 0.5).astype(np.float32)
    arrs = dict(q_db=mk(GQ, H), c_db=mk(GQ * group, H), q_col=mk(GQ, Sq, H), c_col=mk(GQ * group, Sd, H), qL_db=mk(GQ, 3, H),
                cL_db=mk(GQ * group, 3, H), qL_col=mk(GQ, 4, Sq, H), cL_col=mk(GQ * group, 4, Sd, H))
    doc_mask = (np.arange(Sd)[None] < rng.randint(2, Sd + 1, size=GQ * group)[:, None]).astype(np.int64)
    np.savez(tmp_path / "inputs.npz", doc_mask=doc_mask, **arrs)
    world = 2
    mp.spawn(_col_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    t = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in arrs.items()}
    loss, _ = lead_col_distill_loss(t["q_db"], t["c_db"], t["q_col"], t["c_col"], torch.from_numpy(doc_mask), t["qL_db"],
                                    t["cL_db"], t["qL_col"], t["cL_col"], selected_db=[0, 2], selected_col=[1, 3],
                                    temperature=2.0, layer_temperature=0.5, para_dis=1.3, reweight=True, train_col=True,
                                    gather=False)
    loss.backward()
    for r in range(world):
        got = np.load(tmp_path / f"c{r}.npz")
        assert abs(float(got["loss"][0]) - loss.item()) < 1e-5
        for k, v in t.items():
            n = v.shape[0] // world
            assert np.allclose(got[k], v.grad.numpy()[r * n:(r + 1) * n], atol=1e-6, rtol=1e-4), (r, k)
