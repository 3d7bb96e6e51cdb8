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
