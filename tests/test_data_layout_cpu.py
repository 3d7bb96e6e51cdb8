"""CPU: simxns_b200.data (pack_example / collate) against tests/golden/batch_layout.npz — the tensors the reference's own
Rocketqa_v2Dataset.__getitem__ and collate function built (SimANS/utils/MARCO_until_new.py:165-260, run unmodified by
oracle/make_golden.py data with a deterministic tokenizer) from the recorded token lists.  Integer work: bit-exact."""
import numpy as np
import pytest
import torch

from simxns_b200.data import collate, pack_example


def test_pack_and_collate_match_the_reference_bit_exactly(golden_dir):
    g = np.load(golden_dir / "batch_layout.npz")
    feats = []
    saw_truncated_q = saw_full_ctx = False
    for i in range(int(g["n_examples"][0])):
        lens = g[f"ex{i}_ctx_lens"]
        flat = g[f"ex{i}_ctx_tokens"]
        ctx_tokens = np.split(flat, np.cumsum(lens)[:-1])
        q, ctx, ce = pack_example(g[f"ex{i}_q_tokens"].tolist(), [c.tolist() for c in ctx_tokens], pad_token_id=0,
                                  sep_token_id=102)
        assert q.dtype == ctx.dtype == ce.dtype == torch.int64
        assert np.array_equal(q.numpy(), g[f"ex{i}_q"])
        assert np.array_equal(ctx.numpy(), g[f"ex{i}_ctx"])
        assert np.array_equal(ce.numpy(), g[f"ex{i}_ce"])
        saw_truncated_q |= len(g[f"ex{i}_q_tokens"]) == 32
        saw_full_ctx |= bool((lens == 128).any())
        feats.append((q, ctx, ce))
    assert saw_truncated_q and saw_full_ctx  # the fixture exercises the maximum lengths (CE input = 32 + 126 = 158 <= 160)
    batch = collate(feats)
    for k, v in zip(("q", "q_mask", "d", "d_mask"), batch["student"][:4]):
        assert np.array_equal(v.numpy(), g["batch_" + k]), k
    assert batch["student"][4] == g["batch_pos"].tolist()
    for k, v in zip(("ce", "ce_mask", "tgt"), batch["teacher"]):
        assert np.array_equal(v.numpy(), g["batch_" + k]), k
    # layout facts the kernels rely on: passages of query i at rows i*(1+n) .., positive first
    B, group = batch["teacher"][2].shape
    assert batch["student"][2].shape[0] == B * group and batch["student"][4] == [i * group for i in range(B)]
    assert (batch["teacher"][2][:, 0] == 1).all() and (batch["teacher"][2][:, 1:] == 0).all()


def test_pack_example_edge_cases():
    q, ctx, ce = pack_example([101, 7, 102], [[101, 8, 102, 9, 102], [101, 5, 6]], pad_token_id=0, sep_token_id=102)
    assert ce[0, :6].tolist() == [101, 7, 102, 8, 102, 9] and ce[0, 6] == 0   # trailing [SEP] of the passage dropped
    assert ce[1, :5].tolist() == [101, 7, 102, 5, 6] and ce[1, 5] == 0        # no trailing [SEP]: only [CLS] dropped
    assert q.shape == (32,) and ctx.shape == (2, 128) and ce.shape == (2, 160)
    with pytest.raises(ValueError):
        pack_example(list(range(1, 40)), [[101, 5, 102]])


def test_wiki_collate_matches_the_reference_bit_exactly(golden_dir):
    """TraditionDataset.get_collate_fn (SimANS/utils/util_wiki.py:665-698): everything padded to the batch maximum"""
    from simxns_b200.data import collate_wiki

    g = np.load(golden_dir / "batch_layout.npz")
    qs = [a.tolist() for a in np.split(g["wiki_q_flat"], np.cumsum(g["wiki_q_lens"])[:-1])]
    ctxs = [a.tolist() for a in np.split(g["wiki_ctx_flat"], np.cumsum(g["wiki_ctx_lens"])[:-1])]
    feats = []
    for i, q in enumerate(qs):
        c = ctxs[3 * i:3 * i + 3]
        ces = [q + x[1:-1] for x in c]
        feats.append((q, c, ces, [[7, 8]], [[len(q), len(x)] for x in ces]))
    b = collate_wiki(feats)
    for k, v in zip(("ce", "ce_mask", "tgt"), b["reranker"]):
        assert np.array_equal(v.numpy(), g["wiki_" + k]), k
    for k, v in zip(("q", "q_mask", "d", "d_mask"), b["retriever"][:4]):
        assert np.array_equal(v.numpy(), g["wiki_" + k]), k
    assert b["retriever"][4] == g["wiki_pos"].tolist()
    assert np.array_equal(np.array(b["reranker_ctx_start_end"]), g["wiki_start_end"])
    assert b["retriever"][0].shape[1] == max(g["wiki_q_lens"]) and b["retriever"][2].shape[1] == max(g["wiki_ctx_lens"])


def test_collate_int32_option_has_the_same_values():
    import torch
    from simxns_b200.data import collate, pack_example

    feats = [pack_example([101, 7, 8, 102], [[101, 5, 6, 102], [101, 9, 102]]) for _ in range(3)]
    a, b = collate(feats), collate(feats, dtype=torch.int32)
    for key in ("student", "teacher"):
        for x, y in zip(a[key], b[key]):
            if torch.is_tensor(x) and x is not a["teacher"][2]:
                assert y.dtype == torch.int32 and torch.equal(x, y.long())
    assert b["teacher"][2].dtype == torch.int64  # tgt_tensor keeps the reference's dtype (labels, never sent to the kernels)
