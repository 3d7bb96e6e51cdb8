"""B200 (`-m gpu`): exact inner-product top-k (simxns_b200.search.FlatIPIndex over simx_topk_*) against the float64
oracle (oracle/search_numpy.py).  Index work is bit-exact: the returned row numbers must equal the oracle's wherever
the oracle's neighbouring scores are separated by more than the fp32-class score error (1e-5 relative); scores within
2e-6 of the scale; exact ties resolve to the lowest row."""
import numpy as np
import pytest
import torch

from oracle import search_numpy as sn
from simxns_b200 import _lib
from simxns_b200.search import FlatIPIndex, merge_candidates

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    _lib.require_device()


def check(P, Q, k, D, I):
    Dr, Ir, S = sn.flat_ip_search(P, Q, k)
    D, I = D.cpu().numpy(), I.cpu().numpy()
    scale = np.abs(S).max() + 1e-30
    fin = np.isfinite(Dr)
    assert np.array_equal(np.isfinite(D), fin)
    assert np.abs(D[fin] - Dr[fin]).max() < 2e-6 * scale
    assert (I[~fin] == -1).all()
    # rows returned are real, distinct, and in non-increasing score order
    for r in range(Q.shape[0]):
        ids = I[r][I[r] >= 0]
        assert len(set(ids.tolist())) == len(ids)
    assert (np.diff(D, axis=1)[fin[:, 1:]] <= 0).all()
    # index parity wherever the ranking is decided by more than the score error: slot j is pinned when the oracle's
    # j-th score is separated from both neighbours in the full ranking
    gap = 1e-5 * scale
    full = np.sort(S, axis=1)[:, ::-1]
    kk = min(k, P.shape[0])
    g = full[:, :-1] - full[:, 1:] > gap                      # [nq, N-1]: rank j vs rank j+1
    sep_prev = np.concatenate([np.ones((Q.shape[0], 1), bool), g], axis=1)[:, :kk]
    sep_next = np.concatenate([g, np.ones((Q.shape[0], 1), bool)], axis=1)[:, :kk]
    decided = sep_prev & sep_next
    assert decided.mean() > 0.5  # the test must actually pin most slots
    assert np.array_equal(I[:, :kk][decided], Ir[:, :kk][decided])
    # and as a set, modulo undecided boundary slots, the oracle's scores of the returned rows match the returned scores
    got_scores = np.take_along_axis(S, np.maximum(I[:, :kk], 0), axis=1)
    assert np.abs(got_scores - D[:, :kk]).max() < 2e-6 * scale


@pytest.mark.parametrize("N,nq,H,k", [(1000, 7, 64, 10),        # single segment
                                      (50, 3, 128, 200),        # fewer rows than k: -1 / -inf padding
                                      (32768, 4, 64, 200),      # exactly one segment
                                      (40001, 5, 128, 1000),    # two segments, ragged tail, odd N (GEMM padding rows)
                                      (300000, 6, 64, 1000),    # several score chunks + a reduce round with > 1 segment
                                      (140000, 130, 768, 200)]) # bert-base width, more queries than one M tile
def test_flat_ip_search_vs_oracle(N, nq, H, k):
    rng = np.random.RandomState(N % 9973 + k)
    # MS-MARCO-like scores: a shared offset direction puts all inner products in a narrow positive band
    base = rng.standard_normal(H)
    P = (0.5 * rng.standard_normal((N, H)) + base).astype(np.float32)
    Q = (0.5 * rng.standard_normal((nq, H)) + base).astype(np.float32)
    index = FlatIPIndex(H)
    half = N // 3
    index.add(P[:half])           # faiss-style incremental add
    index.add(torch.from_numpy(P[half:]))
    assert index.ntotal == N
    D, I = index.search(Q, k)
    check(P, Q, k, D, I)


def test_exact_ties_take_the_lowest_rows_and_capacity_mode():
    H, N, k = 64, 70000, 16
    rng = np.random.RandomState(0)
    P = rng.standard_normal((N, H)).astype(np.float32)
    P[1000:] = P[np.arange(N - 1000) % 1000]  # every row repeats 70 times: exact score ties everywhere
    Q = rng.standard_normal((3, H)).astype(np.float32)
    index = FlatIPIndex(H, capacity=N)
    for b in range(0, N, 25000):
        index.add(P[b:b + 25000])
    D, I = index.search(Q, k, as_numpy=True)
    Dr, Ir, _ = sn.flat_ip_search(P, Q, k)
    assert np.array_equal(I, Ir)  # bit-exact: identical rows give identical fp32 scores, ties -> ascending row number
    assert np.abs(D - Dr).max() < 2e-6 * np.abs(Dr).max()


def test_sharded_merge_equals_global_search():
    """co.shard=True semantics: two shards with global row offsets, merged candidates == search over the whole corpus"""
    H, N, k = 64, 90000, 100
    rng = np.random.RandomState(3)
    P = rng.standard_normal((N, H)).astype(np.float32)
    Q = rng.standard_normal((4, H)).astype(np.float32)
    cut = 50001
    a, b = FlatIPIndex(H), FlatIPIndex(H, row_offset=cut)
    a.add(P[:cut]); b.add(P[cut:])
    Da, Ia = a.search_local(Q, k)
    Db, Ib = b.search_local(Q, k)
    D, I = merge_candidates(torch.cat([Da, Db], 1), torch.cat([Ia, Ib], 1), k)
    check(P, Q, k, D, I)


def test_bad_arguments_raise():
    with pytest.raises(ValueError):
        FlatIPIndex(100)
    idx = FlatIPIndex(64)
    with pytest.raises(RuntimeError, match="empty"):
        idx.search(np.zeros((1, 64), np.float32), 5)
    idx.add(np.zeros((10, 64), np.float32))
    with pytest.raises(RuntimeError, match="k="):
        idx.search(np.zeros((1, 64), np.float32), 2000)
