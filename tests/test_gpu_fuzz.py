"""B200 (`-m gpu`): seeded random sweep over the tower's configuration space — architecture, depth, width, batch, sequence
length (every kernel variant: S <= 128 single key tile, 128 < S <= 256 two key tiles, longer: blocked), ragged masks, dropout
on/off, gradient checkpointing, int32 / int64 ids, [CLS]-only or full-sequence cotangent — each draw checked forward and
backward against the numpy oracle (oracle/bert_numpy.py, which follows LEAD/modeling_bert.py and modeling_distilbert.py) in
the fp32-class mode at the north_star tolerance (1e-3; asserted at 2e-5 / 3e-4).  The fixed cases in test_gpu_parity.py and
test_gpu_dropout.py pin named edge cases; this file guards the combinations nobody listed."""
import numpy as np
import pytest
import torch

from oracle import bert_numpy as bn
from simxns_b200 import _lib
from test_gpu_parity import X3, assert_grads, build, cu, max_rel

pytestmark = pytest.mark.gpu

N_DRAWS = 14


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    _lib.require_device()


def _draw(i):
    r = np.random.RandomState(9000 + i)
    heads = int(r.choice([1, 2, 3]))
    s_class = i % 4                                       # cycle through the attention kernel variants
    S = int({0: r.randint(1, 33), 1: r.randint(33, 129), 2: r.randint(129, 257), 3: r.randint(257, 400)}[s_class])
    return dict(
        arch=["bert", "distilbert"][(i // 2) % 2], layers=int(r.choice([1, 2, 3])), heads=heads,
        inter=int(r.choice([64, 192, 320])), B=int(r.randint(1, 5)) if S > 128 else int(r.randint(1, 9)), S=S,
        ragged=bool(r.randint(0, 2)), p_hidden=float(r.choice([0.0, 0.1, 0.3])), p_attn=float(r.choice([0.0, 0.1, 0.25])),
        ckpt=bool(r.randint(0, 2)), int32=bool(r.randint(0, 2)), seq_cot=bool(r.randint(0, 2)), seed=int(r.randint(1, 2**31)))


@pytest.mark.parametrize("i", range(N_DRAWS))
def test_random_configuration_vs_oracle(i):
    d = _draw(i)
    cfg = bn.bert_config(num_layers=d["layers"], hidden=64 * d["heads"], heads=d["heads"], intermediate=d["inter"], vocab=211,
                         max_pos=512, arch=d["arch"])
    params = bn.synth_params(cfg, seed=d["seed"] % 1000)
    ids, mask = bn.synth_tokens(d["B"], d["S"], cfg["vocab"], seed=i, ragged=d["ragged"], min_len=1)
    drop = d["p_hidden"] > 0 or d["p_attn"] > 0
    dr = dict(p_hidden=d["p_hidden"], p_attn=d["p_attn"], seed=d["seed"]) if drop else None
    ora = bn.EncoderOracle(cfg, params)
    out = ora.forward(ids, mask, dropout=dr) if drop else ora.forward(ids, mask)
    rng = np.random.RandomState(i)
    d_cls = rng.standard_normal(out["cls"].shape)
    d_seq = rng.standard_normal(out["seq"].shape) * mask[:, :, None] if d["seq_cot"] else None
    ref = ora.backward(d_cls=d_cls, d_seq=d_seq)

    model = build(cfg, params, X3)
    if drop:
        model.hidden_dropout_prob, model.attention_probs_dropout_prob, model.forced_dropout_seed = \
            d["p_hidden"], d["p_attn"], d["seed"]
    model.train()
    if d["ckpt"]:
        model.gradient_checkpointing_enable()
    dt = torch.int32 if d["int32"] else torch.int64
    cls, seq = model.encode(cu(ids).to(dt), cu(mask).to(dt), want_sequence=d["seq_cot"])
    assert max_rel(cls.detach().cpu().numpy(), out["cls"]) < 2e-5, d
    loss = (cls * cu(d_cls, torch.float32)).sum()
    if d["seq_cot"]:
        got = seq.detach().cpu().numpy() * mask[:, :, None]          # padded rows carry no contract
        assert max_rel(got, out["seq"] * mask[:, :, None]) < 2e-5, d
        loss = loss + (seq * cu(d_seq, torch.float32)).sum()
    loss.backward()
    assert_grads(model, ref, 3e-4)
