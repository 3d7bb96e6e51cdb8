"""oracle/ref_torch.py (the CPU-baseline port of the reference wrapper over transformers.BertModel) against the numpy
oracle and the committed golden fixture — i.e. the thing bench.py times on the host computes the reference's numbers."""
import numpy as np
import torch

from oracle import bert_numpy as bn
from oracle import losses_numpy as ln
from oracle import ref_torch as rt


def test_port_matches_oracle_and_golden(golden_dir):
    g = np.load(golden_dir / "encoder_tiny_bert.npz")
    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64)
    params = bn.synth_params(cfg, seed=int(g["seed"][0]))
    enc = rt.make_encoder(cfg)
    rt.load_params(enc, params)
    enc.eval()
    seq, pooled, _ = enc(input_ids=torch.from_numpy(g["ids"]), attention_mask=torch.from_numpy(g["mask"]))
    assert np.abs(seq.detach().numpy() - g["seq"]).max() < 2e-5
    assert np.abs(pooled.detach().numpy() - g["pooled"]).max() < 2e-5


def test_port_losses_match_oracle():
    rng = np.random.RandomState(0)
    Q, Cm = rng.standard_normal((4, 16)).astype(np.float32), rng.standard_normal((12, 16)).astype(np.float32)
    t = rng.standard_normal((4, 3)).astype(np.float32)
    ref, _, _ = ln.score_loss(Q, Cm, terms=["nll_global"], group=3, pos_idx=np.arange(4) * 3)
    assert abs(rt.nll_global_loss(torch.from_numpy(Q), torch.from_numpy(Cm), 3).item() - ref["total"]) < 1e-5
    ref, _, _ = ln.score_loss(Q, Cm, terms=["kl_local"], group=3, teacher=t)
    assert abs(rt.simans_kl_loss(torch.from_numpy(Q), torch.from_numpy(Cm), torch.from_numpy(t)).item() - ref["total"]) < 1e-5
