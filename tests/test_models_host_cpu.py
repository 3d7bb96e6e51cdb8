"""CPU checks of the host-side mirror of the reference interface added in round 2 (no kernel runs): checkpoints with the
old-transformers buffer keys load through every parent module, PROD's role-driven init_encoder picks the
checkpoint and depth the reference picks (PROD/ProD_KD/model/models.py:20-72), dropout / gradient-checkpointing settings
follow the HF config the way the reference's init_encoder writes them, and the per-call C struct carries them."""
import ctypes as C
import json
from types import SimpleNamespace

import pytest
import torch

from simxns_b200 import _lib, lead_models, models, prod_models
from simxns_b200.encoder import EncoderConfig, next_dropout_seed

TINY = dict(num_layers=2, hidden=64, heads=1, intermediate=128, vocab=100, max_pos=32)


def _with_old_buffers(sd, prefixes):
    sd = dict(sd)
    for p in prefixes:
        sd[p + "embeddings.position_ids"] = torch.arange(32).unsqueeze(0)
        sd[p + "embeddings.token_type_ids"] = torch.zeros(1, 32, dtype=torch.long)
    return sd


def test_old_transformers_buffers_are_ignored_under_any_parent_strict():
    enc = lambda: models.HFBertEncoder(EncoderConfig(**TINY))
    bi = models.BiBertEncoder(None, question_model=enc(), ctx_model=enc())
    sd = _with_old_buffers(bi.state_dict(), ["question_model.", "ctx_model."])
    bi2 = models.BiBertEncoder(None, question_model=enc(), ctx_model=enc())
    res = bi2.load_state_dict(sd, strict=True)  # wiki/co_training_wiki_train.py:364 loads strict
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(bi2.ctx_model.embeddings.word_embeddings.weight, bi.ctx_model.embeddings.word_embeddings.weight)
    rr = models.Reranker(enc(), 64)
    rr2 = models.Reranker(enc(), 64)
    rr2.load_state_dict(_with_old_buffers(rr.state_dict(), ["encoder."]), strict=True)
    pr = prod_models.Reranker(enc(), 64)
    prod_models.Reranker(enc(), 64).load_state_dict(_with_old_buffers(pr.state_dict(), ["encoder."]), strict=True)
    alone = enc()
    alone.load_state_dict(_with_old_buffers(alone.state_dict(), [""]), strict=True)
    # a genuinely unexpected key is still an error
    bad = dict(bi.state_dict())
    bad["question_model.embeddings.bogus"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="bogus"):
        bi2.load_state_dict(bad, strict=True)


@pytest.fixture()
def hf_dirs(tmp_path):
    """two local HF checkpoint directories (student: 4 layers hidden 64; teacher: 6 layers hidden 128)"""
    transformers = pytest.importorskip("transformers")
    out = {}
    for name, L, H in (("student", 4, 64), ("teacher", 6, 128)):
        d = tmp_path / name
        d.mkdir()
        cfg = transformers.BertConfig(hidden_size=H, num_hidden_layers=L, num_attention_heads=H // 64,
                                      intermediate_size=2 * H, vocab_size=100, max_position_embeddings=32)
        cfg.save_pretrained(str(d))
        torch.manual_seed(L)
        m = models.HFBertEncoder(EncoderConfig(num_layers=L, hidden=H, heads=H // 64, intermediate=2 * H, vocab=100,
                                               max_pos=32))
        torch.save({"bert." + k: v for k, v in m.state_dict().items()}, str(d / "pytorch_model.bin"))
        out[name] = (str(d), m)
    return out


def test_prod_init_encoder_roles_pick_checkpoint_and_depth(hf_dirs):
    args = SimpleNamespace(model_type=hf_dirs["student"][0], teacher_model_type=hf_dirs["teacher"][0],
                           double_teacher_pretrain=hf_dirs["teacher"][0], student_num_hidden_layers=2,
                           teacher_num_hidden_layers=6, double_teacher_num_hidden_layers=3, gradient_checkpointing=True)
    s = prod_models.HFBertEncoder.init_encoder(args, "student")          # positional role, as the reference drivers call it
    t = prod_models.HFBertEncoder.init_encoder(args, role="teacher")
    dt = prod_models.HFBertEncoder.init_encoder(args, "double_teacher", 0.0)
    assert (s.enc_cfg.num_layers, s.enc_cfg.hidden) == (2, 64)
    assert (t.enc_cfg.num_layers, t.enc_cfg.hidden) == (6, 128)
    assert (dt.enc_cfg.num_layers, dt.enc_cfg.hidden) == (3, 128)
    # weights come from the role's own checkpoint (truncated to the role's depth)
    assert torch.equal(t.encoder.layer[5].output.dense.weight, hf_dirs["teacher"][1].encoder.layer[5].output.dense.weight)
    assert torch.equal(s.encoder.layer[1].output.dense.weight, hf_dirs["student"][1].encoder.layer[1].output.dense.weight)
    assert torch.equal(dt.encoder.layer[2].output.dense.weight, hf_dirs["teacher"][1].encoder.layer[2].output.dense.weight)
    # dropout / checkpointing as init_encoder writes them into the config
    assert (s.hidden_dropout_prob, s.attention_probs_dropout_prob, s.gradient_checkpointing) == (0.1, 0.1, True)
    assert (dt.hidden_dropout_prob, dt.attention_probs_dropout_prob) == (0.1, 0.1)  # dropout=0 keeps the checkpoint's config
    n = prod_models.HFBertEncoder.init_encoder(args, model_type=hf_dirs["teacher"][0], number_layers=4)
    assert n.enc_cfg.num_layers == 4
    with pytest.raises(ValueError, match="no such type role"):
        prod_models.HFBertEncoder.init_encoder(args, "pupil")
    bi = prod_models.BiBertEncoder(args, "student")
    assert bi.role == "student" and bi.ctx_model.enc_cfg.num_layers == 2 and bi.ctx_model is not bi.question_model


def test_simans_init_encoder_dropout_and_checkpointing(hf_dirs):
    args = SimpleNamespace(model_type=hf_dirs["student"][0], gradient_checkpointing=False)
    m = models.HFBertEncoder.init_encoder(args)
    assert (m.hidden_dropout_prob, m.attention_probs_dropout_prob, m.gradient_checkpointing) == (0.1, 0.1, False)
    m2 = models.HFBertEncoder.init_encoder(args, dropout=0.25)
    assert m2.hidden_dropout_prob == 0.25
    m2.gradient_checkpointing_enable()
    assert m2.gradient_checkpointing
    # the per-call struct: probabilities only in train mode, policy from the flag, seed passed through
    c = m2._c_cfg(dropout_seed=123, id_bytes=4)
    assert (round(c.hidden_dropout, 6), round(c.attn_dropout, 6), c.dropout_seed, c.save_policy, c.id_bytes) == \
        (0.25, 0.25, 123, _lib.SAVE_LAYER_INPUTS, 4)
    m2.eval()
    c = m2._c_cfg()
    assert (c.hidden_dropout, c.attn_dropout) == (0.0, 0.0)
    lib = _lib.lib()
    full = m._c_cfg(train=False)
    assert lib.simx_encoder_saved_bytes(C.byref(c), 8, 32) * 4 < lib.simx_encoder_saved_bytes(C.byref(full), 8, 32)
    bad = m._c_cfg(train=False)
    bad.hidden_dropout = 1.5
    assert lib.simx_encoder_saved_bytes(C.byref(bad), 8, 32) == 0 and b"dropout" in lib.simx_last_error()
    bad = m._c_cfg(train=False)
    bad.id_bytes = 2
    assert lib.simx_encoder_workspace_bytes(C.byref(bad), 8, 32) == 0


def test_dropout_seeds_follow_torch_seed_and_advance():
    torch.manual_seed(1234)
    a = [next_dropout_seed() for _ in range(3)]
    assert len(set(a)) == 3 and all(0 <= s < 2 ** 64 for s in a)
    torch.manual_seed(4321)
    assert next_dropout_seed() not in a


def test_distilbert_dropout_fields():
    cfg = SimpleNamespace(model_type="distilbert", n_layers=2, dim=64, n_heads=1, hidden_dim=128, vocab_size=100,
                          max_position_embeddings=32, dropout=0.1, attention_dropout=0.2,
                          hidden_dropout_prob=0.5)  # LEAD sets the BERT-named field, which DistilBERT ignores
    m = lead_models.HFDistilBertEncoder(cfg)
    assert (m.hidden_dropout_prob, m.attention_probs_dropout_prob) == (0.1, 0.2)
