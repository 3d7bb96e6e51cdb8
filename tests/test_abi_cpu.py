"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/simx_b200.h declares,
the ctypes signature table covers exactly those symbols, and the shape/layout queries (no GPU needed) are consistent
with the HF parameter table.  No compute call is made."""
import ctypes as C
import re

import pytest
import torch

from simxns_b200 import _lib
from simxns_b200.encoder import EncoderConfig
from simxns_b200.models import BiBertEncoder, HFBertEncoder, HFDistilBertEncoder, Reranker

from conftest import ROOT


def _declared_symbols():
    text = (ROOT / "include" / "simx_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(simx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/simx_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes SIGNATURES and the header disagree"
    assert lib.simx_version() >= 100


def test_struct_mirrors_have_header_layout():
    assert C.sizeof(_lib.EncoderCfg) == 10 * 4 + 2 * 4 + 8 + 2 * 4  # + dropout probabilities, seed, save_policy, id_bytes
    assert _lib.EncoderCfg.dropout_seed.offset == 48 and _lib.EncoderCfg.id_bytes.offset == 60
    assert C.sizeof(_lib.LossCfg) == 9 * 4 + 14 * 4
    assert C.sizeof(_lib.PairCfg) == 7 * 4 + 11 * 4
    assert C.sizeof(_lib.OptTensor) == 56


@pytest.mark.parametrize("arch,n_expected", [("bert", 5 + 16 * 12 + 2), ("distilbert", 4 + 16 * 6)])
def test_param_table_matches_hf_state_dict(arch, n_expected):
    lib = _lib.lib()
    ec = EncoderConfig(arch=arch, num_layers=12 if arch == "bert" else 6)
    cls = HFBertEncoder if arch == "bert" else HFDistilBertEncoder
    with torch.device("meta"):
        m = cls(ec)
    cfg = ec.c_struct(_lib.PREC_BF16)
    n = lib.simx_encoder_num_params(C.byref(cfg))
    assert n == n_expected == len(m.ordered_parameters()) == len(m.state_dict())
    if arch == "bert":
        assert sum(p.numel() for p in m.parameters()) == 109482240  # params per BERT-base tower
    # every parameter has its slot in the flat gradient buffer, slots tile it exactly
    slots = []
    for i, p in enumerate(m.ordered_parameters()):
        numel = lib.simx_encoder_param_numel(C.byref(cfg), i)
        assert numel == p.numel()
        slots.append((lib.simx_encoder_grad_offset(C.byref(cfg), i), numel))
    slots.sort()
    pos = 0
    for off, numel in slots:
        assert off == pos
        pos += numel
    assert pos == lib.simx_encoder_grad_numel(C.byref(cfg))
    # q/k/v weights (and biases) are adjacent: the fused QKV wgrad writes them as one block
    base = 5 if arch == "bert" else 4
    H = ec.hidden
    oq, ok, ov = (lib.simx_encoder_grad_offset(C.byref(cfg), base + j) for j in (0, 2, 4))
    assert (ok - oq, ov - ok) == (H * H, H * H)


def test_state_dict_keys_are_the_reference_checkpoint_layout():
    """199 HF keys per tower, prefixed question_model. / ctx_model.; Reranker: encoder.* + qa_classifier.*"""
    ec = EncoderConfig(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=64, max_pos=32)
    bi = BiBertEncoder(None, question_model=HFBertEncoder(ec), ctx_model=HFBertEncoder(ec))
    keys = list(bi.state_dict())
    assert len(keys) == 2 * (5 + 16 * 2 + 2)
    assert "question_model.embeddings.word_embeddings.weight" in keys
    assert "ctx_model.encoder.layer.1.attention.output.LayerNorm.bias" in keys
    assert "ctx_model.pooler.dense.weight" in keys
    shared = BiBertEncoder(None, question_model=HFBertEncoder(ec))
    assert shared.ctx_model is shared.question_model
    rr = Reranker(HFBertEncoder(ec), ec.hidden)
    assert {"qa_classifier.weight", "qa_classifier.bias"} <= set(rr.state_dict())
    # the optimizer's no-decay grouping keys on these substrings (co_training_marco_train.py:57-69)
    names = [n for n, _ in bi.named_parameters()]
    assert any("LayerNorm.weight" in n for n in names) and any(n.endswith("bias") for n in names)
    # old checkpoints carry embeddings.position_ids: accepted and ignored
    sd = bi.question_model.state_dict()
    sd["embeddings.position_ids"] = torch.arange(32)[None]
    bi.question_model.load_state_dict(sd, strict=True)


def test_buffer_size_queries_and_invalid_configs():
    lib = _lib.lib()
    ec = EncoderConfig()
    cfg = ec.c_struct(_lib.PREC_BF16)
    per_token = (7 * 768 + 2 * 3072) * 12 + 768 * 14  # saved bf16 elements per token
    saved = lib.simx_encoder_saved_bytes(C.byref(cfg), 2048, 128)
    assert abs(saved - 2 * per_token * 2048 * 128) / saved < 0.01
    cfg3 = ec.c_struct(_lib.PREC_BF16X3)
    assert lib.simx_encoder_saved_bytes(C.byref(cfg3), 4, 32) > 2.5 * lib.simx_encoder_saved_bytes(C.byref(cfg), 4, 32)
    bad = EncoderConfig(hidden=100, heads=2).c_struct(_lib.PREC_BF16)
    assert lib.simx_encoder_num_params(C.byref(bad)) == -1
    assert "hidden" in _lib.last_error()


def test_hidden_state_entry_points_validate_their_arguments():
    """simx_encoder_hidden_state / simx_encoder_add_cotangent reject a depth outside [0, L], a batch stride smaller than one
    sequence and null buffers before touching any memory (checked here with dummy non-null addresses, no GPU involved)"""
    lib = _lib.lib()
    cfg = EncoderConfig(num_layers=2, hidden=64, heads=1, intermediate=64, vocab=32, max_pos=16).c_struct(_lib.PREC_BF16)
    fake = 4096
    assert lib.simx_encoder_hidden_state(C.byref(cfg), fake, 2, 8, 3, fake, 8 * 64, None) != 0
    assert "index 3" in _lib.last_error()
    assert lib.simx_encoder_hidden_state(C.byref(cfg), fake, 2, 8, 1, fake, 8 * 64 - 4, None) != 0
    assert "out_stride_b" in _lib.last_error()
    assert lib.simx_encoder_hidden_state(C.byref(cfg), None, 2, 8, 1, fake, 8 * 64, None) != 0
    assert lib.simx_encoder_add_cotangent(C.byref(cfg), fake, 2, 8, None, 8 * 64, None) != 0
    assert lib.simx_encoder_add_cotangent(C.byref(cfg), fake, 2, 8, fake, 8, None) != 0
    assert "d_stride_b" in _lib.last_error()


def test_forward_without_cuda_fails_loudly():
    """no CPU fallback: parameters on the CPU are an error, not a slow path"""
    ec = EncoderConfig(num_layers=1, hidden=64, heads=1, intermediate=64, vocab=32, max_pos=16)
    m = HFBertEncoder(ec)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m(input_ids=torch.ones(1, 4, dtype=torch.long), attention_mask=torch.ones(1, 4, dtype=torch.long))


def test_widened_rows_refuse_cpu_tensors_and_bad_arguments():
    """no CPU fallback anywhere: the optimizer, the clip, the index and the MSE helper raise on CPU tensors"""
    from simxns_b200 import optim as sx
    from simxns_b200.losses import embed_mse
    from simxns_b200.search import FlatIPIndex

    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sx.AdamW([p]).step()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sx.clip_grad_norm_([p], 1.0)
    with pytest.raises(NotImplementedError):
        sx.clip_grad_norm_([p], 1.0, norm_type=1.0)
    with pytest.raises(ValueError):
        sx.AdamW([p], flavor="lion")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        embed_mse(torch.zeros(2, 8), torch.zeros(2, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FlatIPIndex(64, device="cpu")
    with pytest.raises(ValueError):
        FlatIPIndex(100, device="cuda:0")


def test_opt_plan_chunks_the_tensor_table():
    """simx_opt_plan is host-only: 8192-element chunks, prefix in chunk_begin, -1 on a bad row"""
    lib = _lib.lib()
    rows = (_lib.OptTensor * 3)()
    for i, n in enumerate((1, 8192, 8193)):
        rows[i].p, rows[i].numel = 16 * (i + 1), n
    assert lib.simx_opt_plan(rows, 3) == 1 + 1 + 2
    assert [rows[i].chunk_begin for i in range(3)] == [0, 1, 2]
    rows[1].numel = 0
    assert lib.simx_opt_plan(rows, 3) == -1 and b"tensor 1" in lib.simx_last_error()


def test_linear_schedule_and_optimizer_groups_host_logic():
    """get_linear_schedule_with_warmup values and get_optimizer's bias / LayerNorm.weight no-decay split
    (co_training_marco_train.py:57-69, 129-134) — pure host logic"""
    import types

    from simxns_b200.train import get_linear_schedule_with_warmup, get_optimizer

    ec = EncoderConfig(num_layers=1, hidden=64, heads=1, intermediate=128, vocab=50, max_pos=16)
    m = HFBertEncoder(ec)
    opt = get_optimizer(types.SimpleNamespace(optimizer="adamW"), m, weight_decay=0.01, lr=2e-5, eps=1e-8)
    decay, no_decay = opt.param_groups
    names = {id(p): n for n, p in m.named_parameters()}
    assert all("bias" in names[id(p)] or "LayerNorm.weight" in names[id(p)] for p in no_decay["params"])
    assert not any("bias" in names[id(p)] or "LayerNorm.weight" in names[id(p)] for p in decay["params"])
    assert decay["weight_decay"] == 0.01 and no_decay["weight_decay"] == 0.0
    assert len(decay["params"]) + len(no_decay["params"]) == len(list(m.parameters()))
    with pytest.raises(Exception, match="not recognized"):
        get_optimizer(types.SimpleNamespace(optimizer="sgd"), m)
    sched = get_linear_schedule_with_warmup(opt, num_warmup_steps=2, num_training_steps=10)
    lrs = []
    for _ in range(11):
        lrs.append(opt.param_groups[0]["lr"])
        sched.step()  # (stepping the scheduler alone is fine for reading the schedule)
    want = [2e-5 * (k / 2 if k < 2 else max(0.0, (10 - k) / 8)) for k in range(11)]
    assert all(abs(a - b) < 1e-12 for a, b in zip(lrs, want))


def test_merge_candidates_orders_by_score_then_id():
    from simxns_b200.search import merge_candidates

    D = torch.tensor([[1.0, 3.0, 3.0, float("-inf"), 2.0, 3.0]])
    I = torch.tensor([[7, 9, 4, -1, 5, 6]])
    d, i = merge_candidates(D, I, 4)
    assert i.tolist() == [[4, 6, 9, 5]] and d.tolist() == [[3.0, 3.0, 3.0, 2.0]]
    d, i = merge_candidates(D, I, 6)
    assert i.tolist()[0][-1] == -1 and d.tolist()[0][-1] == float("-inf")


def test_roberta_modules_have_the_reference_key_sets():
    """HFRobertaEncoder == HF RobertaModel keys (with the dead pooler); RobertaDot == roberta.* without pooler +
    embeddingHead / norm (SimANS/model/models.py:168-192, 333-352); the two pooler slots of the kernel's parameter
    table are non-persistent zero buffers there"""
    from simxns_b200.models import HFRobertaEncoder, RobertaDot

    ec = EncoderConfig(num_layers=1, hidden=64, heads=1, intermediate=128, vocab=50, max_pos=20, type_vocab=1, ln_eps=1e-5,
                       pad_token_id=1, roberta_positions=True)
    enc = HFRobertaEncoder(ec)
    assert "pooler.dense.weight" in enc.state_dict() and len(enc.state_dict()) == 5 + 16 + 2
    assert enc.embeddings.word_embeddings.padding_idx == 1 and enc.embeddings.position_embeddings.padding_idx == 1
    dot = RobertaDot(ec)
    keys = list(dot.state_dict())
    assert not any("pooler" in k for k in keys) and len(keys) == 5 + 16 + 4
    assert {"embeddingHead.weight", "embeddingHead.bias", "norm.weight", "norm.bias"} <= set(keys)
    assert len(dot.roberta.ordered_parameters()) == 5 + 16 + 2  # the table still has the two pooler slots
    assert ec.pooler is True  # the caller's config object is not modified
