"""B200 (`-m gpu`): every depth's FULL hidden state as a differentiable output of the tower — HF `output_hidden_states`,
LEAD's `all_layer_hidden [B, L, S, H]` (LEAD/models.py:125-138, :184-196) — through simx_encoder_hidden_state (forward)
and simx_encoder_add_cotangent between the layer slices of the backward pass.

  * against the numpy oracle with a random cotangent on EVERY depth (plus [CLS] rows and the sequence output), both
    architectures, dropout, checkpointing, both gradient-delivery modes;
  * against the reference's own HFColBertEncoder (tests/golden/lead_colbert_tiny.npz, oracle/make_golden.py): token
    embeddings, per-layer states, colbert_score on the output and on an intermediate layer, and the parameter gradients
    of the resulting loss."""
import numpy as np
import pytest
import torch

from oracle import bert_numpy as bn
from simxns_b200 import _lib
from simxns_b200 import lead_models
from simxns_b200.losses import colbert_score
from test_gpu_parity import X3, BF16, assert_grads, build, cu, ec_from, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    _lib.require_device()


@pytest.mark.parametrize("arch,B,S,drop,ckpt,delivery", [
    ("bert", 3, 40, False, False, "tower"), ("bert", 2, 130, True, False, "layer"), ("distilbert", 4, 24, True, True, "tower"),
    ("bert", 1, 300, False, True, "layer"), ("distilbert", 2, 77, False, False, "layer")])
def test_hidden_states_forward_backward_vs_oracle(arch, B, S, drop, ckpt, delivery):
    cfg = bn.bert_config(num_layers=3, hidden=128, heads=2, intermediate=192, vocab=300, max_pos=512, arch=arch)
    params = bn.synth_params(cfg, seed=5 + S)
    ids, mask = bn.synth_tokens(B, S, cfg["vocab"], seed=S, ragged=True, min_len=3)
    dr = dict(p_hidden=0.1, p_attn=0.1, seed=424242) if drop else None
    ora = bn.EncoderOracle(cfg, params)
    out = ora.forward(ids, mask, dropout=dr) if drop else ora.forward(ids, mask)
    want_hidden = np.stack(out["hidden"], axis=1)                      # [B, L+1, S, H]
    rng = np.random.RandomState(S)
    d_hidden = rng.standard_normal(want_hidden.shape) * mask[:, None, :, None]
    d_cls = rng.standard_normal(out["cls"].shape)
    d_seq = rng.standard_normal(out["seq"].shape) * mask[:, :, None]
    ref = ora.backward(d_cls=d_cls, d_seq=d_seq, d_hidden=d_hidden)
    m4 = mask[:, None, :, None]
    for precision, tol_h, tol_g in ((X3, 2e-5, 3e-4), (BF16, 6e-2, 1.2e-1)):
        model = build(cfg, params, precision)
        model.grad_delivery = delivery
        if drop:
            model.hidden_dropout_prob, model.attention_probs_dropout_prob, model.forced_dropout_seed = 0.1, 0.1, 424242
        model.train()
        if ckpt:
            model.gradient_checkpointing_enable()
        cls, seq, hidden = model.encode(cu(ids), cu(mask), want_sequence=True, want_hidden=True)
        assert tuple(hidden.shape) == want_hidden.shape
        assert max_rel(hidden.detach().cpu().numpy() * m4, want_hidden * m4) < tol_h, (precision, "hidden")
        assert torch.equal(hidden[:, -1], seq) and torch.equal(hidden[:, :, 0], cls)
        loss = (hidden * cu(d_hidden, torch.float32)).sum() + (cls * cu(d_cls, torch.float32)).sum() + \
            (seq * cu(d_seq, torch.float32)).sum()
        loss.backward()
        assert_grads(model, ref, tol_g)
        # the same states without autograd (a frozen teacher): only the layer inputs are kept during that forward
        with torch.no_grad():
            _, _, h2 = model.encode(cu(ids), cu(mask), want_hidden=True)
        assert torch.equal(h2, hidden)


def test_only_an_intermediate_depth_carries_gradient():
    """no cotangent on the outputs at all — the pass is seeded by hidden state 1 alone; layers above it get exact zeros"""
    cfg = bn.bert_config(num_layers=3, hidden=64, heads=1, intermediate=128, vocab=200, max_pos=64)
    params = bn.synth_params(cfg, seed=2)
    ids, mask = bn.synth_tokens(3, 20, cfg["vocab"], seed=8, ragged=True)
    ora = bn.EncoderOracle(cfg, params)
    out = ora.forward(ids, mask)
    d_hidden = np.zeros((3, 4, 20, 64))
    d_hidden[:, 1] = np.random.RandomState(0).standard_normal((3, 20, 64)) * mask[:, :, None]
    ref = ora.backward(d_hidden=d_hidden)
    model = build(cfg, params, X3).train()
    _, _, hidden = model.encode(cu(ids), cu(mask), want_hidden=True)
    (hidden[:, 1] * cu(d_hidden[:, 1], torch.float32)).sum().backward()
    zero_true = tuple(n for n, _ in model.named_parameters() if "layer.1." in n or "layer.2." in n or "pooler" in n)
    assert_grads(model, ref, 3e-4, zero_true=zero_true)
    for n, p in model.named_parameters():
        if n in zero_true:
            assert float(p.grad.abs().max()) == 0.0, n


def test_colbert_tower_vs_reference_golden(golden_dir):
    g = np.load(golden_dir / "lead_colbert_tiny.npz")
    cfg = bn.bert_config(num_layers=3, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64)
    params = bn.synth_params(cfg, seed=int(g["seed"][0]))
    layer = int(g["layer"][0])
    for precision, tol_h, tol_s, tol_g in ((X3, 2e-5, 1e-4, 3e-4), (BF16, 5e-2, 5e-2, 1.5e-1)):
        model = lead_models.HFColBertEncoder(ec_from(cfg), precision=precision)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
        model = model.cuda().eval()
        q_out, q_hid, _ = model("colbert_query", "cuda", input_ids=cu(g["q_ids"]), attention_mask=cu(g["q_mask"]))
        d_out, d_hid, _ = model("colbert_doc", "cuda", input_ids=cu(g["d_ids"]), attention_mask=cu(g["d_mask"]))
        assert tuple(q_hid.shape) == g["q_hid"].shape and tuple(d_hid.shape) == g["d_hid"].shape
        mq, md = g["q_mask"][:, None, :, None], g["d_mask"][:, None, :, None]
        assert max_rel(q_hid.detach().cpu().numpy() * mq, g["q_hid"] * mq) < tol_h
        assert max_rel(d_hid.detach().cpu().numpy() * md, g["d_hid"] * md) < tol_h
        assert max_rel(d_out.detach().cpu().numpy() * md[:, 0], g["d_out"] * md[:, 0]) < tol_h
        # padded QUERY positions take part in the reference's sum (util.py:459-465), so their states matter too
        assert max_rel(q_out.detach().cpu().numpy(), g["q_out"]) < tol_h
        scores = colbert_score(q_out, d_out, cu(g["d_mask"]), 2)
        layer_scores = colbert_score(q_hid[:, layer], d_hid[:, layer], cu(g["d_mask"]), 2) / 64.0
        assert max_rel(scores.detach().cpu().numpy(), g["scores"]) < tol_s
        assert max_rel(layer_scores.detach().cpu().numpy(), g["layer_scores"]) < tol_s
        target = torch.tensor([0, 3], device="cuda")
        lsm = torch.nn.functional.log_softmax
        loss = torch.nn.functional.nll_loss(lsm(scores, dim=-1), target) + \
            0.5 * torch.nn.functional.nll_loss(lsm(layer_scores, dim=-1), target)
        assert abs(loss.item() - float(g["loss"][0])) < (1e-4 if precision == X3 else 5e-2)
        loss.backward()
        ref = {k[5:]: g[k] for k in g.files if k.startswith("grad/")}
        assert_grads(model, ref, tol_g, zero_true=("pooler.dense.weight", "pooler.dense.bias"))


def test_virt_attention_distillation_through_both_towers():
    """DE teacher (BERT) -> DB student (DistilBERT) with the token-level attention-map loss (--distill_de_db_attention,
    LEAD/util.py:617-618, :643-644, :766-769): LEAD's towers with full_layer_hidden, losses.query_doc_attention_map /
    virt_loss on the CUDA tensors; expected gradients = the same loss evaluated on the oracle's hidden states (float64) and
    pushed through the oracle's backward."""
    import types

    from simxns_b200.losses import query_doc_attention_map, virt_loss

    Q, D, Sq, Sd = 2, 4, 10, 20
    args = types.SimpleNamespace(device="cuda", temperature=1.0)
    sel_s, sel_t = [0, 1], [1, 3]
    towers = {}
    for name, arch, layers, seed in (("t", "bert", 4, 61), ("s", "distilbert", 2, 62)):
        cfg = bn.bert_config(num_layers=layers, hidden=128, heads=2, intermediate=192, vocab=300, max_pos=64, arch=arch)
        towers[name] = (cfg, bn.synth_params(cfg, seed=seed))
    q_ids, q_mask = bn.synth_tokens(Q, Sq, 300, seed=63, ragged=True, min_len=3)
    d_ids, d_mask = bn.synth_tokens(D, Sd, 300, seed=64, ragged=True, min_len=5)
    qlen, dlen = q_mask.sum(1).tolist(), d_mask.sum(1).tolist()

    # oracle side: hidden states in float64 -> torch leaf tensors -> the same loss -> d_hidden -> oracle backward
    want, leaves, oracles = {}, {}, {}
    for name, (cfg, params) in towers.items():
        for side, ids, mask in (("q", q_ids, q_mask), ("d", d_ids, d_mask)):
            o = bn.EncoderOracle(cfg, params)
            h = np.stack(o.forward(ids, mask)["hidden"], axis=1)
            oracles[name, side] = o
            leaves[name, side] = torch.from_numpy(h).requires_grad_(True)
    a64 = types.SimpleNamespace(device="cpu", temperature=1.0)
    mt, _, _ = query_doc_attention_map(a64, leaves["t", "q"][:, 1:], leaves["t", "d"][:, 1:], qlen, dlen, sel_t)
    ms, _, _ = query_doc_attention_map(a64, leaves["s", "q"][:, 1:], leaves["s", "d"][:, 1:], qlen, dlen, sel_s)
    loss64 = virt_loss(mt, ms, 3.0)
    loss64.backward()
    for name in towers:
        gq = oracles[name, "q"].backward(d_hidden=leaves[name, "q"].grad.numpy())
        gd = oracles[name, "d"].backward(d_hidden=leaves[name, "d"].grad.numpy())
        want[name] = {k: gq[k] + gd[k] for k in gq}           # one shared tower encodes queries and documents

    models = {}
    for name, (cfg, params) in towers.items():
        cls_ = lead_models.HFBertEncoder if cfg["arch"] == "bert" else lead_models.HFDistilBertEncoder
        m = cls_(ec_from(cfg), precision=X3)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
        m.full_layer_hidden = True
        models[name] = m.cuda().eval()
    hid = {}
    for name, m in models.items():
        hid[name, "q"] = m(input_ids=cu(q_ids), attention_mask=cu(q_mask))[4]
        hid[name, "d"] = m(input_ids=cu(d_ids), attention_mask=cu(d_mask))[4]
        assert tuple(hid[name, "d"].shape) == (D, towers[name][0]["num_layers"], Sd, 128)
    mt, _, _ = query_doc_attention_map(args, hid["t", "q"], hid["t", "d"], qlen, dlen, sel_t)
    ms, _, _ = query_doc_attention_map(args, hid["s", "q"], hid["s", "d"], qlen, dlen, sel_s)
    loss = virt_loss(mt, ms, 3.0)
    assert abs(loss.item() - loss64.item()) < 1e-4 * max(1.0, abs(loss64.item()))
    loss.backward()
    for name, m in models.items():
        top = towers[name][0]["num_layers"] - 1
        dead = tuple(n for n, _ in m.named_parameters() if "pooler" in n or
                     any(f"layer.{l}." in n for l in range(max(sel_t if name == "t" else sel_s) + 1, top + 1)))
        assert_grads(m, want[name], 1e-3, zero_true=dead)   # north_star gradient tolerance (the maps exponentiate |logits| ~ 30)


def test_get_loss_dual_dense_branch_is_the_fused_global_kernel():
    """LEAD/util.py:416-451 for dual encoders: log-softmax over all passages, positives at i * (1+n); loss and gradients
    against the same formula in torch (float64)"""
    import types

    from simxns_b200.losses import get_loss_dual

    g = torch.Generator().manual_seed(5)
    q = torch.randn(6, 128, generator=g).cuda().requires_grad_(True)
    c = torch.randn(24, 128, generator=g).cuda().requires_grad_(True)
    loss = get_loss_dual(types.SimpleNamespace(model_type="dual_encoder"), q, c, None)
    loss.backward()
    q64, c64 = q.detach().double().requires_grad_(True), c.detach().double().requires_grad_(True)
    want = torch.nn.functional.nll_loss(torch.log_softmax(q64 @ c64.t(), dim=1), torch.arange(0, 24, 4, device="cuda"))
    want.backward()
    assert abs(loss.item() - want.item()) < 1e-4
    assert torch.allclose(q.grad.double(), q64.grad, atol=1e-4, rtol=1e-3)
    assert torch.allclose(c.grad.double(), c64.grad, atol=1e-4, rtol=1e-3)


def test_pass_state_is_released_by_reference_counting():
    """outputs -> grad_fn -> pass state must not loop back to the outputs: with the cyclic collector switched off the memory
    in use after each step (saved activations, hidden states, the flat gradient buffer) has to return to the same level"""
    import gc

    cfg = bn.bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=128)
    model = build(cfg, bn.synth_params(cfg, seed=1), BF16).train()
    ids, mask = bn.synth_tokens(64, 128, cfg["vocab"], seed=3)
    ids, mask = cu(ids), cu(mask)
    gc.collect()
    gc.disable()
    try:
        levels = []
        for _ in range(4):
            cls, seq, hidden = model.encode(ids, mask, want_sequence=True, want_hidden=True)
            (cls.sum() + seq.sum() + hidden[:, 1].sum()).backward()
            model.zero_grad(set_to_none=True)
            del cls, seq, hidden
            model.last_cls_layers = None
            torch.cuda.synchronize()
            levels.append(torch.cuda.memory_allocated())
        assert levels[1] == levels[2] == levels[3], levels
    finally:
        gc.enable()
