"""CPU checks of the dropout generator's numpy restatement (oracle/dropout_numpy.py): the Philox4x32 round function against
the published Random123 known-answer vectors, the keep probability / scaling of the masks, and independence across rows,
columns, layers, sites and seeds.  The CUDA generator (csrc/dropout.cuh) is checked against this one on the GPU
(tests/test_gpu_dropout.py)."""
import numpy as np
import pytest

from oracle import bert_numpy as bn
from oracle import dropout_numpy as dn


def _u32(xs):
    return [int(x, 16) for x in xs.split()]


@pytest.mark.parametrize("ctr,key,want", [
    ("00000000 00000000 00000000 00000000", "00000000 00000000", "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
    ("ffffffff ffffffff ffffffff ffffffff", "ffffffff ffffffff", "408f276d 41c83b0e a20bc7c6 6d5451fd"),
    ("243f6a88 85a308d3 13198a2e 03707344", "a4093822 299f31d0", "d16cfe09 94fdcceb 5001e420 24126ea1"),
])
def test_philox4x32_10_known_answers(ctr, key, want):
    """Random123 kat_vectors, philox4x32 with 10 rounds: pins the round function, multipliers and key schedule that the
    7-round product generator shares"""
    c, k = _u32(ctr), _u32(key)
    got = dn.philox7(c[0], c[1], c[2], c[3], k[0], k[1], rounds=10)
    assert [int(np.asarray(g).reshape(-1)[0]) for g in got] == _u32(want)


def test_threshold_and_scale_are_unbiased():
    for p in (0.1, 0.25, 0.5):
        thr = dn.threshold(p)
        assert abs(thr / 65536 - p) < 1e-5
        assert abs(dn.keep_scale(p) * (1 - thr / 65536) - 1.0) < 1e-12


def test_hidden_mask_statistics():
    p = 0.1
    m = dn.hidden_mask(seed=0x1234567890ABCDEF, site=dn.SITE_FFN_OUT, layer=3, T=2048, H=768, p=p)
    keep = (m > 0)
    n = keep.size
    assert abs(keep.mean() - (1 - p)) < 4 * np.sqrt(p * (1 - p) / n)      # keep rate
    assert abs(m.mean() - 1.0) < 4 * np.sqrt(p / (1 - p) / n)             # unbiased after scaling
    assert np.allclose(np.unique(m), [0.0, dn.keep_scale(p)])
    # rows and columns are balanced too (no structure along either axis)
    assert np.abs(keep.mean(0) - (1 - p)).max() < 6 * np.sqrt(p * (1 - p) / keep.shape[0])
    assert np.abs(keep.mean(1) - (1 - p)).max() < 6 * np.sqrt(p * (1 - p) / keep.shape[1])
    # neighbouring elements (the two halves of one Philox word, adjacent words, adjacent calls) are uncorrelated
    k = keep.astype(np.float64) - keep.mean()
    for shift in (1, 2, 8):
        c = (k[:, :-shift] * k[:, shift:]).mean() / k.var()
        assert abs(c) < 5 / np.sqrt(n)
    c = (k[:-1] * k[1:]).mean() / k.var()
    assert abs(c) < 5 / np.sqrt(n)


def test_masks_differ_across_seed_site_layer_plane():
    base = dict(seed=42, site=dn.SITE_ATTN_OUT, layer=0, T=64, H=128, p=0.5)
    m0 = dn.hidden_mask(**base)
    assert np.array_equal(m0, dn.hidden_mask(**base))  # a pure function of its arguments
    for change in (dict(seed=43), dict(site=dn.SITE_FFN_OUT), dict(layer=1), dict(seed=42 + (1 << 32))):
        m1 = dn.hidden_mask(**{**base, **change})
        agree = ((m0 > 0) == (m1 > 0)).mean()
        assert 0.4 < agree < 0.6, (change, agree)
    a = dn.attn_mask(seed=42, layer=0, B=2, heads=2, S=24, p=0.5)
    assert a.shape == (2, 2, 24, 24)
    for x, y in ((a[0, 0], a[0, 1]), (a[0, 0], a[1, 0])):
        assert 0.35 < ((x > 0) == (y > 0)).mean() < 0.65


def test_oracle_dropout_is_identity_at_p0_and_unbiased_in_expectation():
    cfg = bn.bert_config(num_layers=1, hidden=64, heads=1, intermediate=128, vocab=128, max_pos=16)
    P = bn.synth_params(cfg, seed=0)
    ids, mask = bn.synth_tokens(4, 12, 128, seed=1)
    ref = bn.EncoderOracle(cfg, P).forward(ids, mask)["cls"]
    same = bn.EncoderOracle(cfg, P).forward(ids, mask, dropout=dict(p_hidden=0.0, p_attn=0.0, seed=5))["cls"]
    assert np.array_equal(ref, same)
    drop = bn.EncoderOracle(cfg, P).forward(ids, mask, dropout=dict(p_hidden=0.1, p_attn=0.1, seed=5))["cls"]
    assert not np.allclose(ref, drop)
    assert np.abs(drop).max() < 10 * np.abs(ref).max()
