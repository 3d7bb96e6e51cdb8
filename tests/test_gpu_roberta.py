"""B200 (`-m gpu`): the RoBERTa towers of SimANS MS-Doc (simxns_b200.models.HFRobertaEncoder / RobertaDot over
simx_encoder_forward_pos / simx_encoder_backward_pos) against tests/golden/roberta_tiny.npz — outputs and gradients of the
reference's own classes (SimANS/model/models.py:168-192, 333-352) — in the fp32-class mode at the north_star tolerances,
and the bf16 mode at the bf16 tolerances."""
import numpy as np
import pytest
import torch

from oracle import bert_numpy as bn
from simxns_b200 import _lib
from simxns_b200.encoder import EncoderConfig
from simxns_b200.models import HFRobertaEncoder, RobertaDot

pytestmark = pytest.mark.gpu
X3, BF16 = _lib.PREC_BF16X3, _lib.PREC_BF16


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    _lib.require_device()


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def roberta_ec(pooler=True):
    return EncoderConfig(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=66, type_vocab=1, ln_eps=1e-5,
                         pad_token_id=1, roberta_positions=True, pooler=pooler)


def check_grads(named, ref, tol, skip=()):
    scales = [np.abs(v).max() for v in ref.values() if np.abs(v).max() > 0]
    floor = 1e-3 * float(np.median(scales))
    for name, p in named:
        if name in skip:
            continue
        assert p.grad is not None, name
        r = ref[name]
        zero_true = "key.bias" in name
        e = float(np.abs(p.grad.cpu().numpy() - r).max() / max(np.abs(r).max(), max(scales) if zero_true else floor))
        assert e < tol, (name, e)


@pytest.mark.parametrize("precision,tol_h,tol_g", [(X3, 2e-5, 1e-3), (BF16, 5e-2, 1e-1)])
def test_hf_roberta_encoder_vs_reference_golden(golden_dir, precision, tol_h, tol_g):
    g = np.load(golden_dir / "roberta_tiny.npz")
    cfg = bn.roberta_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=66)
    enc = HFRobertaEncoder(roberta_ec(), precision=precision)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in bn.synth_params(cfg, seed=int(g["seed"][0])).items()}, strict=True)
    enc = enc.cuda()
    seq, pooled, none = enc(input_ids=cu(g["ids"]), attention_mask=cu(g["mask"]))
    assert none is None
    valid = g["mask"].astype(bool)
    sc = np.abs(g["enc_seq"][valid]).max()
    assert np.abs(seq.detach().cpu().numpy()[valid] - g["enc_seq"][valid]).max() < tol_h * sc
    assert np.abs(pooled.detach().cpu().numpy() - g["enc_pooled"]).max() < tol_h * sc
    (seq * cu(g["enc_d_seq"])).sum().backward()
    ref = {k[len("enc_grad/"):]: g[k] for k in g.files if k.startswith("enc_grad/")}
    check_grads(enc.named_parameters(), ref, tol_g)
    # nn.Embedding(padding_idx=1) on both tables: the pad rows get exactly nothing
    assert float(enc.embeddings.word_embeddings.weight.grad[1].abs().max()) == 0.0
    assert float(enc.embeddings.position_embeddings.weight.grad[1].abs().max()) == 0.0
    assert float(enc.pooler.dense.weight.grad.abs().max()) == 0.0  # dead pooler: dense zeros, not None


def test_roberta_dot_vs_reference_golden(golden_dir):
    g = np.load(golden_dir / "roberta_tiny.npz")
    cfg = bn.roberta_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=66)
    dot = RobertaDot(roberta_ec(), precision=X3)
    params = bn.synth_params(cfg, seed=int(g["seed"][0]))
    sd = {"roberta." + k: torch.from_numpy(v) for k, v in params.items() if not k.startswith("pooler.")}
    sd.update({"embeddingHead.weight": torch.from_numpy(g["head_w"]), "embeddingHead.bias": torch.from_numpy(g["head_b"]),
               "norm.weight": torch.from_numpy(g["norm_w"]), "norm.bias": torch.from_numpy(g["norm_b"])})
    dot.load_state_dict(sd, strict=True)  # exactly the reference's key set (no pooler: add_pooling_layer=False)
    dot = dot.cuda()
    emb = dot(cu(g["ids"]), cu(g["mask"]), True)
    assert np.abs(emb.detach().cpu().numpy() - g["dot_emb"]).max() < 1e-3
    assert torch.equal(emb, dot.body_emb(cu(g["ids"]), cu(g["mask"])))
    (emb * cu(g["dot_d_emb"])).sum().backward()
    ref = {k[len("dot_grad/"):]: g[k] for k in g.files if k.startswith("dot_grad/")}
    assert sorted(n for n, _ in dot.named_parameters()) == sorted(ref)
    check_grads(dot.named_parameters(), ref, 1e-3)


def test_roberta_position_limit_raises():
    enc = HFRobertaEncoder(roberta_ec(), precision=X3).cuda()
    ids = torch.full((1, 65), 5, dtype=torch.long, device="cuda")
    with pytest.raises(ValueError, match="max_position_embeddings"):
        enc(input_ids=ids, attention_mask=torch.ones_like(ids))
