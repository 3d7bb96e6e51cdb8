"""CPU: the oracle's RoBERTa mode (position ids from padding, pad id 1 on both embedding tables) against
tests/golden/roberta_tiny.npz — outputs and gradients of the reference's own HFRobertaEncoder and RobertaDot
(SimANS/model/models.py:168-192, 333-352; oracle/make_golden.py roberta)."""
import numpy as np

from oracle import bert_numpy as bn


def _close(a, b, atol, rtol=2e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= atol + rtol * np.abs(b).max(), np.abs(a - b).max()


def test_position_ids_follow_hf_roberta():
    cfg = bn.roberta_config(num_layers=1, hidden=64, heads=1, intermediate=64, vocab=50, max_pos=20)
    ids = np.array([[0, 5, 6, 2, 1, 1], [0, 7, 2, 1, 1, 1], [0, 2, 1, 1, 1, 1]])
    assert bn.position_ids(cfg, ids).tolist() == [[2, 3, 4, 5, 1, 1], [2, 3, 4, 1, 1, 1], [2, 3, 1, 1, 1, 1]]
    assert bn.position_ids(bn.bert_config(), ids).tolist() == [[0, 1, 2, 3, 4, 5]] * 3


def test_hf_roberta_encoder_and_roberta_dot(golden_dir):
    g = np.load(golden_dir / "roberta_tiny.npz")
    cfg = bn.roberta_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=66)
    params = bn.synth_params(cfg, seed=int(g["seed"][0]))
    enc = bn.EncoderOracle(cfg, params)
    out = enc.forward(g["ids"], g["mask"])
    valid = g["mask"].astype(bool)
    _close(out["seq"][valid], g["enc_seq"][valid], 2e-5)
    _close(out["cls"][:, -1], g["enc_pooled"], 2e-5)
    grads = enc.backward(d_seq=g["enc_d_seq"].astype(np.float64))
    n = 0
    for name in bn.param_names(cfg):
        ref = g["enc_grad/" + name]
        _close(grads[name], ref, 1e-5 * max(1.0, np.abs(ref).max()))
        n += 1
    assert n == 5 + 16 * 2 + 2
    assert np.all(grads["embeddings.word_embeddings.weight"][1] == 0) and np.all(grads["embeddings.position_embeddings.weight"][1] == 0)
    # RobertaDot: norm(embeddingHead(sequence_output[:, 0])) and its gradients into the head and the tower
    cls = out["cls"][:, -1]
    pre = cls @ g["head_w"].astype(np.float64).T + g["head_b"]
    emb, cache = bn.layer_norm(pre, g["norm_w"].astype(np.float64), g["norm_b"].astype(np.float64), 1e-5)
    _close(emb, g["dot_emb"], 2e-5)
    dpre, dnw, dnb = bn.layer_norm_bwd(g["dot_d_emb"].astype(np.float64), g["norm_w"].astype(np.float64), cache)
    _close(dnw, g["dot_grad/norm.weight"], 1e-5)
    _close(dnb, g["dot_grad/norm.bias"], 1e-5)
    _close(dpre.T @ cls, g["dot_grad/embeddingHead.weight"], 1e-5)
    _close(dpre.sum(0), g["dot_grad/embeddingHead.bias"], 1e-5)
    d_cls = np.zeros(out["cls"].shape)
    d_cls[:, -1] = dpre @ g["head_w"].astype(np.float64)
    grads = enc.backward(d_cls=d_cls)
    for name in bn.param_names(cfg):
        if name.startswith("pooler."):
            continue  # RobertaModel(add_pooling_layer=False)
        ref = g["dot_grad/roberta." + name]
        _close(grads[name], ref, 1e-5 * max(1.0, np.abs(ref).max()))
