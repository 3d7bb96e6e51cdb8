"""ORACLE tooling — generates tests/golden/*.npz by running THE REFERENCE'S OWN CODE (imported unmodified from
/root/reference, CPU fp32) on seeded synthetic weights and inputs.  Run in the build container only
(`python oracle/make_golden.py`); the fixtures are committed because /root/reference does not exist on the GPU box.

What is executed from the reference:
  * SimANS/model/models.py        HFBertEncoder, BiBertEncoder, Reranker, BiEncoderNllLoss      (imported)
  * PROD/ProD_KD/model/models.py  CrossBERTKDLoss                                                (imported)
  * LEAD/models.py                HFBertEncoder / HFDistilBertEncoder over the vendored modeling files (imported via the
                                  10-line compatibility shim of SURVEY §8c)
  * the SimANS retriever-step loss is inline in train() (co_training_marco_train.py:199-214) and cannot be imported:
    those lines are re-typed below verbatim in `simans_kl_loss`.
Weights come from oracle.bert_numpy.synth_params (seeded numpy), loaded through load_state_dict, so the fixture
does not depend on HF's init order.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"

from oracle.bert_numpy import bert_config, param_names, synth_params, synth_tokens  # noqa: E402


def import_from(project_dir: Path, module: str):
    """import `module` with `project_dir` first on sys.path (the reference's scripts do the same: co_training_marco_train.py:3-11)"""
    for m in list(sys.modules):
        if m == "model" or m.startswith("model.") or m in ("models", "modeling_bert", "modeling_distilbert"):
            del sys.modules[m]
    sys.path.insert(0, str(project_dir))
    try:
        return __import__(module, fromlist=["*"])
    finally:
        sys.path.pop(0)


def hf_bert_config(cfg):
    from transformers import BertConfig

    return BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["num_layers"],
                      num_attention_heads=cfg["heads"], intermediate_size=cfg["intermediate"],
                      max_position_embeddings=cfg["max_pos"], type_vocab_size=cfg["type_vocab"],
                      hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=cfg["ln_eps"])


def load_params(module, params, prefix=""):
    sd = {prefix + k: torch.from_numpy(v.copy()) for k, v in params.items()}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m or "token_type_ids" in m for m in missing), missing


def simans_kl_loss(local_q_vector, local_ctx_vectors, relevance_logits, temperature_distill, eps=1e-7):
    # --- verbatim from /root/reference/SimANS/co_training/co_training_marco_train.py:199-214 (scale_simmila off) ---
    student_local_ctx_vectors = local_ctx_vectors.reshape(local_q_vector.size(0),
                                                          local_ctx_vectors.size(0) // local_q_vector.size(0), -1)
    student_simila = torch.einsum("bh,bdh->bd", local_q_vector, student_local_ctx_vectors)
    student_dist_p = F.softmax(student_simila, dim=1)
    teacher_logits = relevance_logits / temperature_distill
    probs = F.softmax(teacher_logits, dim=1)
    teacher_dist_p = probs
    loss_fct = torch.nn.KLDivLoss(reduction='batchmean')
    distill_loss = loss_fct((student_dist_p + eps).log(), teacher_dist_p)
    return distill_loss


def grads_digest(named_grads):
    """compact per-tensor fingerprint: [l2 norm, sum, abs-max] + first 16 elements"""
    out = {}
    for n, g in named_grads.items():
        g = g.detach().double().reshape(-1)
        out["gstat/" + n] = np.array([g.norm().item(), g.sum().item(), g.abs().max().item()])
        out["ghead/" + n] = g[:16].numpy().copy()
    return out


# -----------------------------------------------------------------------------------------------------
def golden_encoder_tiny(simans):
    """tiny BERT through the reference HFBertEncoder: per-layer CLS, last hidden, ALL gradients under a random cotangent"""
    cfg = bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64)
    params = synth_params(cfg, seed=7)
    ids, mask = synth_tokens(3, 24, cfg["vocab"], seed=11, ragged=True)
    enc = simans.HFBertEncoder(hf_bert_config(cfg))
    load_params(enc, params)
    enc.eval()
    out = enc.forward(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask),
                      output_hidden_states=True) if False else None
    # the wrapper drops hidden states; call it for (sequence_output, pooled_output) and BertModel.forward for layers
    seq, pooled, _ = enc(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask))
    full = super(simans.HFBertEncoder, enc).forward(input_ids=torch.from_numpy(ids),
                                                    attention_mask=torch.from_numpy(mask), output_hidden_states=True)
    cls = torch.stack([h[:, 0, :] for h in full.hidden_states], dim=1)  # [B, L+1, H]
    rng = np.random.RandomState(5)
    d_cls = rng.standard_normal(cls.shape).astype(np.float32)
    d_seq = rng.standard_normal(seq.shape).astype(np.float32)
    (cls * torch.from_numpy(d_cls)).sum().add((seq * torch.from_numpy(d_seq)).sum()).backward()
    res = dict(ids=ids, mask=mask, cls=cls.detach().numpy(), seq=seq.detach().numpy(), pooled=pooled.detach().numpy(),
               d_cls=d_cls, d_seq=d_seq, cfg=np.array([2, 128, 2, 256, 512, 64, 2]), seed=np.array([7]))
    for n, p in enc.named_parameters():
        res["grad/" + n] = p.grad.numpy().copy()
    np.savez_compressed(GOLD / "encoder_tiny_bert.npz", **res)
    print("encoder_tiny_bert", {k: v.shape for k, v in res.items() if not k.startswith("grad/")})


def golden_biencoder_cfg1(simans):
    """BASELINE.json configs[0]: BiBertEncoder bert-base, B=2, q 32 / p 128, 1 pos + 1 hard-neg, SimANS KL loss, fp32 CPU"""
    cfg = bert_config()
    pq = synth_params(cfg, seed=0, perturb=False)
    pc = synth_params(cfg, seed=1, perturb=False)
    q_ids, q_mask = synth_tokens(2, 32, cfg["vocab"], seed=1234, ragged=True)
    p_ids, p_mask = synth_tokens(4, 128, cfg["vocab"], seed=1235, ragged=True, min_len=16)
    teacher = (3.0 * np.random.RandomState(99).standard_normal((2, 2))).astype(np.float32)
    model = simans.BiBertEncoder.__new__(simans.BiBertEncoder)
    torch.nn.Module.__init__(model)
    model.question_model = simans.HFBertEncoder(hf_bert_config(cfg))
    model.ctx_model = simans.HFBertEncoder(hf_bert_config(cfg))
    load_params(model.question_model, pq)
    load_params(model.ctx_model, pc)
    model.eval()
    q, a = model(torch.from_numpy(q_ids), torch.from_numpy(q_mask), torch.from_numpy(p_ids), torch.from_numpy(p_mask))
    loss = simans_kl_loss(q, a, torch.from_numpy(teacher), temperature_distill=1.0)
    loss.backward()
    res = dict(q_ids=q_ids, q_mask=q_mask, p_ids=p_ids, p_mask=p_mask, teacher=teacher, q_embs=q.detach().numpy(),
               a_embs=a.detach().numpy(), loss=np.array([loss.item()]), seeds=np.array([0, 1]))
    res.update(grads_digest({n: p.grad for n, p in model.named_parameters()}))
    # a few complete small tensors
    for n, p in model.named_parameters():
        if n.endswith("encoder.layer.0.attention.self.query.bias") or n.endswith("encoder.layer.11.output.LayerNorm.weight") \
                or n.endswith("embeddings.LayerNorm.bias") or n.endswith("encoder.layer.5.intermediate.dense.bias"):
            res["grad/" + n] = p.grad.numpy().copy()
    np.savez_compressed(GOLD / "biencoder_cfg1.npz", **res)
    print("biencoder_cfg1 loss", loss.item(), "q", q.shape, "a", a.shape)


def golden_biencoder_cfg1_bf16_autocast(simans):
    """The same cfg1 step run by the reference's own classes under torch.autocast(bfloat16) — the reference's reduced-precision
    behaviour (SimANS drivers: apex O1 / --fp16; the B200 configs of BASELINE.json name bf16).  Stored next to the fp32
    fixture so the bf16 throughput mode of the kernels can be held to the deviation the reference itself shows."""
    cfg = bert_config()
    pq = synth_params(cfg, seed=0, perturb=False)
    pc = synth_params(cfg, seed=1, perturb=False)
    q_ids, q_mask = synth_tokens(2, 32, cfg["vocab"], seed=1234, ragged=True)
    p_ids, p_mask = synth_tokens(4, 128, cfg["vocab"], seed=1235, ragged=True, min_len=16)
    teacher = (3.0 * np.random.RandomState(99).standard_normal((2, 2))).astype(np.float32)
    model = simans.BiBertEncoder.__new__(simans.BiBertEncoder)
    torch.nn.Module.__init__(model)
    model.question_model = simans.HFBertEncoder(hf_bert_config(cfg))
    model.ctx_model = simans.HFBertEncoder(hf_bert_config(cfg))
    load_params(model.question_model, pq)
    load_params(model.ctx_model, pc)
    model.eval()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        q, a = model(torch.from_numpy(q_ids), torch.from_numpy(q_mask), torch.from_numpy(p_ids), torch.from_numpy(p_mask))
    loss = simans_kl_loss(q.float(), a.float(), torch.from_numpy(teacher), temperature_distill=1.0)
    loss.backward()
    res = dict(q_embs=q.detach().float().numpy(), a_embs=a.detach().float().numpy(), loss=np.array([loss.item()]))
    res.update(grads_digest({n: p.grad for n, p in model.named_parameters()}))
    for n, p in model.named_parameters():
        if n.endswith("encoder.layer.0.attention.self.query.bias") or n.endswith("encoder.layer.11.output.LayerNorm.weight") \
                or n.endswith("embeddings.LayerNorm.bias") or n.endswith("encoder.layer.5.intermediate.dense.bias"):
            res["grad/" + n] = p.grad.float().numpy().copy()
    np.savez_compressed(GOLD / "biencoder_cfg1_bf16_autocast.npz", **res)
    print("biencoder_cfg1_bf16_autocast loss", loss.item())


def golden_reranker_tiny(simans):
    cfg = bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64)
    params = synth_params(cfg, seed=21)
    ids, mask = synth_tokens(6, 40, cfg["vocab"], seed=22, ragged=True, min_len=8)
    rr = simans.Reranker(simans.HFBertEncoder(hf_bert_config(cfg)), cfg["hidden"])
    load_params(rr.encoder, params)
    rng = np.random.RandomState(23)
    w = (0.05 * rng.standard_normal((1, cfg["hidden"]))).astype(np.float32)
    b = np.array([0.1], np.float32)
    rr.qa_classifier.weight.data = torch.from_numpy(w.copy())
    rr.qa_classifier.bias.data = torch.from_numpy(b.copy())
    rr.eval()
    logits = rr(torch.from_numpy(ids).view(2, 3, 40), torch.from_numpy(mask).view(2, 3, 40))
    # co_training_marco_train.py:229-242 (teacher step): CE against column 0
    loss = torch.nn.CrossEntropyLoss()(logits, torch.zeros(2, dtype=torch.long))
    loss.backward()
    res = dict(ids=ids.reshape(2, 3, 40), mask=mask.reshape(2, 3, 40), qa_w=w, qa_b=b, logits=logits.detach().numpy(),
               loss=np.array([loss.item()]), seed=np.array([21]))
    for n, p in rr.named_parameters():
        res["grad/" + n] = p.grad.numpy().copy()
    np.savez_compressed(GOLD / "reranker_tiny.npz", **res)
    print("reranker_tiny", logits.detach().numpy())


def golden_losses(simans, prod):
    rng = np.random.RandomState(3)
    GQ, group, H = 6, 4, 32
    Q = rng.standard_normal((GQ, H)).astype(np.float32)
    Cm = rng.standard_normal((GQ * group, H)).astype(np.float32)
    teacher = (3 * rng.standard_normal((GQ, group))).astype(np.float32)
    res = dict(Q=Q, C=Cm, teacher=teacher, group=np.array([group]))

    def fresh():
        return torch.from_numpy(Q.copy()).requires_grad_(True), torch.from_numpy(Cm.copy()).requires_grad_(True)

    # (1) BiEncoderNllLoss.calc on the full matrix, positives at i*group  (SimANS/model/models.py:468-505)
    q, c = fresh()
    pos = [i * group for i in range(GQ)]
    loss, correct = simans.BiEncoderNllLoss().calc(q, c, pos)
    loss.backward()
    res.update(nll_loss=np.array([loss.item()]), nll_correct=np.array([int(correct)]), nll_dQ=q.grad.numpy().copy(),
               nll_dC=c.grad.numpy().copy(), nll_pos=np.array(pos))
    # (2) SimANS KL (co_training_marco_train.py:199-214), T = 1 and T = 2
    for T in (1.0, 2.0):
        q, c = fresh()
        loss = simans_kl_loss(q, c, torch.from_numpy(teacher), T)
        loss.backward()
        tag = f"kl_T{int(T)}"
        res.update({tag + "_loss": np.array([loss.item()]), tag + "_dQ": q.grad.numpy().copy(),
                    tag + "_dC": c.grad.numpy().copy()})
    # (3) PROD CrossBERTKDLoss.calc, KD_softmax, CE 0.1 / KD 0.9 / T 4   (PROD/ProD_KD/model/models.py:669-781)
    args = types.SimpleNamespace(KD_type="KD_softmax", TEMPERATURE=4.0, CE_WEIGHT=0.1, KD_WEIGHT=0.9, LwF_WEIGHT=1.0,
                                 DKD_alpha=None, DKD_beta=None)
    q, c = fresh()
    loss, correct = prod.CrossBERTKDLoss().calc(args, q, c, torch.from_numpy(teacher))
    loss.backward()
    res.update(prod_loss=np.array([loss.item()]), prod_correct=np.array([int(correct)]), prod_dQ=q.grad.numpy().copy(),
               prod_dC=c.grad.numpy().copy())
    # (4) PROD prob_loss == MASTER/wiki CE-to-teacher form
    args.KD_type = "prob_loss"
    args.TEMPERATURE = 1.0
    q, c = fresh()
    loss, _ = prod.CrossBERTKDLoss().calc(args, q, c, torch.from_numpy(teacher))
    loss.backward()
    res.update(prob_loss=np.array([loss.item()]), prob_dQ=q.grad.numpy().copy(), prob_dC=c.grad.numpy().copy())
    # (6) SimANS wiki retriever loss with the adversarial term, lines re-typed from wiki/co_training_wiki_train.py:198-228
    for lam in (0.0, 0.3):
        q, c = fresh()
        eps_ = 1e-7
        relevance_logits = torch.from_numpy(teacher)
        retriever_local_ctx_vectors = c.reshape(q.size(0), c.size(0) // q.size(0), -1)
        retriever_simila = torch.einsum("bh,bdh->bd", q, retriever_local_ctx_vectors)
        retriever_dist_p = F.softmax(retriever_simila, dim=1)
        reranker_dist_p = F.softmax(relevance_logits / 2.0, dim=1)
        positive_logits = relevance_logits[:, :1]
        negtive_logits = relevance_logits[:, :]
        positive_logits_expand = positive_logits.expand(negtive_logits.size())
        reward_logits = torch.stack((positive_logits_expand, negtive_logits), -1)
        reward_prob = F.softmax(reward_logits, dim=2)
        reward = torch.log(reward_prob[:, :, 0] + eps_)
        normal_loss = -reranker_dist_p * torch.log(retriever_dist_p + eps_)
        normal_loss = normal_loss.sum() / retriever_dist_p.size(0)
        adv_loss = reward * torch.log(retriever_dist_p + eps_)
        adv_loss = adv_loss.sum()
        loss = lam * adv_loss + (1 - lam) * normal_loss
        loss.backward()
        tag = f"wiki_l{int(lam * 10)}"
        res.update({tag + "_loss": np.array([loss.item()]), tag + "_dQ": q.grad.numpy().copy(), tag + "_dC": c.grad.numpy().copy()})
    # (5) PROD BiEncoderKDLoss.calc: DE -> DE distillation on the global matrix (models.py:970-1059)
    Qt = rng.standard_normal((GQ, H)).astype(np.float32)
    Ct = rng.standard_normal((GQ * group, H)).astype(np.float32)
    args.KD_type, args.TEMPERATURE = "KD_softmax", 4.0
    q, c = fresh()
    loss, correct = prod.BiEncoderKDLoss().calc(args, q, c, torch.from_numpy(Qt), torch.from_numpy(Ct), pos)
    loss.backward()
    res.update(Qt=Qt, Ct=Ct, bikd_loss=np.array([loss.item()]), bikd_correct=np.array([int(correct)]),
               bikd_dQ=q.grad.numpy().copy(), bikd_dC=c.grad.numpy().copy())
    np.savez_compressed(GOLD / "losses.npz", **res)
    print("losses", {k: float(v[0]) for k, v in res.items() if k.endswith("loss")})


def golden_losses_prod_kd(prod):
    """the remaining KD_type branches of PROD's loss classes (PROD/ProD_KD/model/models.py:669-781, 970-1094):
    CrossBERTKDLoss {KD_logit, DKD, KD_softmax + LwF, KD_logit + LwF} and BiEncoderKDLoss {KD_logit, DKD, Bi_logit}"""
    rng = np.random.RandomState(11)
    GQ, group, H = 6, 4, 32
    Q = rng.standard_normal((GQ, H)).astype(np.float32)
    Cm = rng.standard_normal((GQ * group, H)).astype(np.float32)
    Qo = (Q + 0.3 * rng.standard_normal((GQ, H))).astype(np.float32)   # the frozen initial student of LwF
    Co = (Cm + 0.3 * rng.standard_normal((GQ * group, H))).astype(np.float32)
    Qt = rng.standard_normal((GQ, H)).astype(np.float32)
    Ct = rng.standard_normal((GQ * group, H)).astype(np.float32)
    teacher = (3 * rng.standard_normal((GQ, group))).astype(np.float32)
    pos = [i * group for i in range(GQ)]
    res = dict(Q=Q, C=Cm, Qo=Qo, Co=Co, Qt=Qt, Ct=Ct, teacher=teacher, group=np.array([group]), pos=np.array(pos),
               hyper=np.array([0.1, 0.9, 4.0, 1.0, 1.0, 8.0]))  # CE_W, KD_W, T, LwF_W, DKD_alpha, DKD_beta
    args = types.SimpleNamespace(KD_type="KD_logit", TEMPERATURE=4.0, CE_WEIGHT=0.1, KD_WEIGHT=0.9, LwF_WEIGHT=1.0,
                                 DKD_alpha=1.0, DKD_beta=8.0)

    def fresh():
        return torch.from_numpy(Q.copy()).requires_grad_(True), torch.from_numpy(Cm.copy()).requires_grad_(True)

    def put(tag, loss, correct, q, c):
        loss.backward()
        res.update({tag + "_loss": np.array([loss.item()]), tag + "_correct": np.array([int(correct)]),
                    tag + "_dQ": q.grad.numpy().copy(), tag + "_dC": c.grad.numpy().copy()})

    for kd in ("KD_logit", "DKD"):
        args.KD_type = kd
        q, c = fresh()
        put("cross_" + kd, *prod.CrossBERTKDLoss().calc(args, q, c, torch.from_numpy(teacher)), q, c)
        q, c = fresh()
        put("bi_" + kd, *prod.BiEncoderKDLoss().calc(args, q, c, torch.from_numpy(Qt), torch.from_numpy(Ct), pos), q, c)
    for kd in ("KD_softmax", "KD_logit"):
        args.KD_type = kd
        q, c = fresh()
        put("lwf_" + kd, *prod.CrossBERTKDLoss().calc(args, q, c, torch.from_numpy(teacher), LwF=True,
                                                     ori_q_vector=torch.from_numpy(Qo), ori_ctx_vectors=torch.from_numpy(Co)),
            q, c)
    args.KD_type = "Bi_logit"
    q, c = fresh()
    put("bi_Bi_logit", *prod.BiEncoderKDLoss().calc(args, q, c, torch.from_numpy(Qt), torch.from_numpy(Ct), pos), q, c)
    np.savez_compressed(GOLD / "losses_prod_kd.npz", **res)
    print("losses_prod_kd", {k: float(v[0]) for k, v in res.items() if k.endswith("loss")})


def roberta_tokens(B, S, vocab, seed, pad=1, min_len=3):
    """right-padded RoBERTa-style rows: <s>=0 first, </s>=2 last valid, pad=1; mask = ids != pad"""
    rng = np.random.RandomState(seed)
    ids = rng.randint(3, vocab, size=(B, S)).astype(np.int64)
    ids[:, 0] = 0
    lens = rng.randint(min_len, S + 1, size=B)
    for b in range(B):
        ids[b, lens[b] - 1] = 2
        ids[b, lens[b]:] = pad
    return ids, (ids != pad).astype(np.int64)


def golden_roberta(simans):
    """SimANS MS-Doc towers (SimANS/model/models.py): HFRobertaEncoder :168-192 (RobertaModel with the dead pooler) and
    RobertaDot :333-352 (RobertaModel without pooler + embeddingHead + LayerNorm on the first token), run unmodified on
    CPU fp32 with seeded weights; position ids come from HF's create_position_ids_from_input_ids (pad id 1)."""
    from transformers import RobertaConfig

    from oracle.bert_numpy import roberta_config

    cfg = roberta_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=300, max_pos=66)
    hf = RobertaConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["num_layers"],
                       num_attention_heads=cfg["heads"], intermediate_size=cfg["intermediate"],
                       max_position_embeddings=cfg["max_pos"], type_vocab_size=1, hidden_dropout_prob=0.0,
                       attention_probs_dropout_prob=0.0, layer_norm_eps=cfg["ln_eps"], pad_token_id=1)
    import copy

    hf_dot = copy.deepcopy(hf)
    hf_dot.return_dict = False  # what RobertaDot.__init__ sets under transformers 4 (models.py:337-338)
    params = synth_params(cfg, seed=31)
    ids, mask = roberta_tokens(5, 40, cfg["vocab"], seed=32)
    rng = np.random.RandomState(33)
    res = dict(ids=ids, mask=mask, seed=np.array([31]))
    # (1) HFRobertaEncoder: (sequence_output, pooled = sequence_output[:, 0], None)
    enc = simans.HFRobertaEncoder(hf)
    load_params(enc, params)
    enc.eval()
    seq, pooled, _ = enc(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask))
    d_seq = (rng.standard_normal(seq.shape) * mask[:, :, None]).astype(np.float32)
    (seq * torch.from_numpy(d_seq)).sum().backward()
    res.update(enc_seq=seq.detach().numpy(), enc_pooled=pooled.detach().numpy(), enc_d_seq=d_seq)
    for n, p in enc.named_parameters():
        res["enc_grad/" + n] = p.grad.numpy().copy()
    # (2) RobertaDot.query_emb: norm(embeddingHead(first token))
    dot = simans.RobertaDot(hf_dot)
    load_params(dot.roberta, {k: v for k, v in params.items() if not k.startswith("pooler.")})
    head_w = (0.05 * rng.standard_normal((cfg["hidden"], cfg["hidden"]))).astype(np.float32)
    head_b = (0.02 * rng.standard_normal(cfg["hidden"])).astype(np.float32)
    norm_w = (1 + 0.1 * rng.standard_normal(cfg["hidden"])).astype(np.float32)
    norm_b = (0.02 * rng.standard_normal(cfg["hidden"])).astype(np.float32)
    with torch.no_grad():
        dot.embeddingHead.weight.copy_(torch.from_numpy(head_w)); dot.embeddingHead.bias.copy_(torch.from_numpy(head_b))
        dot.norm.weight.copy_(torch.from_numpy(norm_w)); dot.norm.bias.copy_(torch.from_numpy(norm_b))
    dot.eval()
    emb = dot.query_emb(torch.from_numpy(ids), torch.from_numpy(mask))
    d_emb = rng.standard_normal(emb.shape).astype(np.float32)
    (emb * torch.from_numpy(d_emb)).sum().backward()
    res.update(dot_emb=emb.detach().numpy(), dot_d_emb=d_emb, head_w=head_w, head_b=head_b, norm_w=norm_w, norm_b=norm_b)
    for n, p in dot.named_parameters():
        res["dot_grad/" + n] = p.grad.numpy().copy()
    np.savez_compressed(GOLD / "roberta_tiny.npz", **res)
    print("roberta_tiny", float(np.abs(res["enc_pooled"]).max()), float(np.abs(res["dot_emb"]).max()))


def golden_batch_layout():
    """the reference's own Rocketqa_v2Dataset.__getitem__ and collate (SimANS/utils/MARCO_until_new.py:165-260) run on a
    tiny synthetic corpus with a deterministic whitespace tokenizer: records the token lists the tokenizer returned (the
    inputs of simxns_b200.data.pack_example) and the tensors the reference built from them"""
    import random
    import tempfile

    sys.path.insert(0, str(REF / "SimANS"))
    for m in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
        del sys.modules[m]
    from utils.MARCO_until_new import Rocketqa_v2Dataset

    class Tok:
        pad_token_id, sep_token_id, cls_token_id = 0, 102, 101

        def __init__(self):
            self.calls = []

        def _ids(self, text):
            return [1000 + (sum(ord(ch) for ch in w) % 700) for w in text.split()]

        def encode(self, text, text_pair=None, add_special_tokens=True, max_length=128, truncation=True,
                   pad_to_max_length=False):
            a = self._ids(text)
            if text_pair is None:
                ids = [self.cls_token_id] + a[:max_length - 2] + [self.sep_token_id]
            else:
                b = self._ids(text_pair)
                while len(a) + len(b) + 3 > max_length:  # longest_first
                    (a if len(a) > len(b) else b).pop()
                ids = [self.cls_token_id] + a + [self.sep_token_id] + b + [self.sep_token_id]
            self.calls.append(ids)
            return ids

    rng = np.random.RandomState(5)
    words = ["w%d" % i for i in range(400)]
    sent = lambda n: " ".join(rng.choice(words, n))
    tmp = Path(tempfile.mkdtemp())
    n_pass = 40
    with open(tmp / "para.txt", "w") as f:
        for i in range(n_pass):
            f.write(f"{i}\t{sent(rng.randint(5, 200))}\n")  # some passages overflow 128 tokens and are truncated
    with open(tmp / "para.title.txt", "w") as f:
        for i in range(0, n_pass, 2):  # half of the passages have no title -> '-'
            f.write(f"{i}\t{sent(rng.randint(1, 6))}\n")
    with open(tmp / "train.tsv", "w") as f:
        for q in range(4):
            pos = f"{q} {75 + q}"
            negs = ",".join(f"{p} {74.5 - 0.3 * j}" for j, p in enumerate(range(10, 30)))
            f.write(f"{q}\t{sent(rng.randint(3, 50))}\t{pos}\t{negs}\n")  # one query overflows 32 tokens
    tok = Tok()
    ds = Rocketqa_v2Dataset(str(tmp / "train.tsv"), tok, num_hard_negatives=3, corpus_path=str(tmp))
    random.seed(0)
    feats, res = [], {}
    for i in range(len(ds)):
        tok.calls = []
        q, ctx, ce = ds[i]
        feats.append((q, ctx, ce))
        ctx_lists, q_list = tok.calls[:-1], tok.calls[-1]  # __getitem__ tokenises the passages first, then the query
        res[f"ex{i}_q_tokens"] = np.array(q_list)
        res[f"ex{i}_ctx_lens"] = np.array([len(c) for c in ctx_lists])
        res[f"ex{i}_ctx_tokens"] = np.concatenate([np.array(c) for c in ctx_lists])
        res[f"ex{i}_q"], res[f"ex{i}_ctx"], res[f"ex{i}_ce"] = q.numpy(), ctx.numpy(), ce.numpy()
    batch = Rocketqa_v2Dataset.get_collate_fn(None)(feats)
    for k, v in zip(("q", "q_mask", "d", "d_mask"), batch["student"][:4]):
        res["batch_" + k] = v.numpy()
    res["batch_pos"] = np.array(batch["student"][4])
    for k, v in zip(("ce", "ce_mask", "tgt"), batch["teacher"]):
        res["batch_" + k] = v.numpy()
    res["n_examples"] = np.array([len(ds)])
    # the NQ / TriviaQA collate (SimANS/utils/util_wiki.py:665-698) on ragged synthetic features: pads to the batch maximum
    from utils.util_wiki import TraditionDataset

    wrng = np.random.RandomState(9)
    wfeats = []
    for i in range(3):
        q = [101] + wrng.randint(1000, 2000, wrng.randint(3, 40)).tolist() + [102]
        ctxs = [[101] + wrng.randint(1000, 2000, wrng.randint(5, 120)).tolist() + [102] for _ in range(3)]
        ces = [q + c[1:-1] for c in ctxs]
        wfeats.append((q, ctxs, ces, [[7, 8]], [[len(q), len(c)] for c in ces]))
    wb = TraditionDataset.get_collate_fn(None)(wfeats)
    res["wiki_q_flat"] = np.concatenate([np.array(f[0]) for f in wfeats])
    res["wiki_q_lens"] = np.array([len(f[0]) for f in wfeats])
    res["wiki_ctx_flat"] = np.concatenate([np.array(c) for f in wfeats for c in f[1]])
    res["wiki_ctx_lens"] = np.array([len(c) for f in wfeats for c in f[1]])
    for k, v in zip(("ce", "ce_mask", "tgt"), wb["reranker"]):
        res["wiki_" + k] = v.numpy()
    for k, v in zip(("q", "q_mask", "d", "d_mask"), wb["retriever"][:4]):
        res["wiki_" + k] = v.numpy()
    res["wiki_pos"] = np.array(wb["retriever"][4])
    res["wiki_start_end"] = np.array(wb["reranker_ctx_start_end"])
    np.savez_compressed(GOLD / "batch_layout.npz", **res)
    print("batch_layout", {k: v.shape for k, v in res.items() if k.startswith("batch_")})


def golden_lead_select_layer():
    """LEAD/util.py:206-254 select_layer, the reference function itself, over every branch (seeded `random`)"""
    import contextlib
    import io
    import json
    import random

    lead_shim()
    util = import_from(REF / "LEAD", "util")
    cases = []
    for mode in ("random", "last", "skip"):
        for name in ("distilbert-base-uncased", "bert-large-uncased"):
            for add_linear in (False, True):
                for train_file in ("data/ms/train.tsv", "data/nq/train.json"):
                    for k in (3, 5):
                        a = dict(layer_selection_random=mode == "random", layer_selection_last=mode == "last",
                                 layer_selection_skip=mode == "skip", pretrained_model_name=name, add_linear=add_linear,
                                 train_file=train_file, disitll_layer_num=k)
                        random.seed(len(cases))
                        with contextlib.redirect_stdout(io.StringIO()):
                            db, te = util.select_layer(types.SimpleNamespace(**a))
                        cases.append(dict(args=a, seed=len(cases), db=db, teacher=te))
    (GOLD / "lead_select_layer.json").write_text(json.dumps(cases))
    print("lead_select_layer", len(cases), "cases")


def lead_shim():
    """make LEAD's vendored HF-4.x modeling files importable under transformers 5"""
    import transformers
    import transformers.pytorch_utils as pu

    if not hasattr(pu, "find_pruneable_heads_and_indices"):
        pu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    if "transformers.deepspeed" not in sys.modules:
        m = types.ModuleType("transformers.deepspeed")
        m.is_deepspeed_zero3_enabled = lambda: False
        sys.modules["transformers.deepspeed"] = m
    return transformers


def golden_lead_distilbert():
    """tiny DistilBERT through LEAD's vendored modeling_distilbert (the 6-layer student arch of config 5)"""
    lead_shim()
    try:
        md = import_from(REF / "LEAD", "modeling_distilbert")
    except Exception as e:  # pragma: no cover
        print("LEAD modeling_distilbert not importable:", repr(e))
        return
    from transformers import DistilBertConfig

    cfg = bert_config(num_layers=2, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64, arch="distilbert")
    params = synth_params(cfg, seed=31)
    ids, mask = synth_tokens(3, 24, cfg["vocab"], seed=32, ragged=True)
    hcfg = DistilBertConfig(vocab_size=512, max_position_embeddings=64, n_layers=2, n_heads=2, dim=128, hidden_dim=256,
                            dropout=0.0, attention_dropout=0.0, output_hidden_states=True)
    if not hasattr(md.DistilBertPreTrainedModel, "get_head_mask"):
        md.DistilBertPreTrainedModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    model = md.DistilBertModel(hcfg)
    load_params(model, params)
    model.eval()
    # vendored forward returns (BaseModelOutput, all_hidden_states, all_hidden_states_all_head); the hidden-state
    # tuple has one entry per layer and EXCLUDES the embedding output (LEAD/modeling_distilbert.py:344-376)
    result, hs, _ = model(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask),
                          output_hidden_states=True, return_dict=True)
    hs = list(hs)
    seq = result.last_hidden_state
    res = dict(ids=ids, mask=mask, seq=seq.detach().numpy(), n_hidden=np.array([len(hs)]), seed=np.array([31]),
               hidden_cls=np.stack([h[:, 0, :].detach().numpy() for h in hs], axis=1))
    rng = np.random.RandomState(33)
    d_seq = rng.standard_normal(seq.shape).astype(np.float32)
    (seq * torch.from_numpy(d_seq)).sum().backward()
    res["d_seq"] = d_seq
    for n, p in model.named_parameters():
        res["grad/" + n] = p.grad.numpy().copy()
    np.savez_compressed(GOLD / "encoder_tiny_distilbert.npz", **res)
    print("encoder_tiny_distilbert hidden entries:", len(hs))


def golden_lead_distill():
    """the real LEAD distill_loss (DE teacher -> DB student, both trained, per-layer score distillation with
    re-weighting) on synthetic embeddings / per-layer [CLS] rows: loss + every gradient"""
    lead_shim()
    util = import_from(REF / "LEAD", "util")
    rng = np.random.RandomState(41)
    GQ, group, H, k = 6, 4, 32, 3
    mk = lambda *shape: (rng.standard_normal(shape) * 0.5).astype(np.float32)
    arrs = dict(q_db=mk(GQ, H), c_db=mk(GQ * group, H), q_de=mk(GQ, H), c_de=mk(GQ * group, H),
                qL_db=mk(GQ, 6, 1, H), cL_db=mk(GQ * group, 6, 1, H), qL_de=mk(GQ, 12, 1, H), cL_de=mk(GQ * group, 12, 1, H))
    sel_db, sel_de = [1, 3, 5], [7, 9, 11]
    res = dict(arrs, sel_db=np.array(sel_db), sel_de=np.array(sel_de), group=np.array([group]))
    for tag, train_de, reweight, T in (("both", True, True, 2.0), ("frozen", False, False, 1.0)):
        t = {n: torch.from_numpy(v.copy()).requires_grad_(True) for n, v in arrs.items()}
        args = types.SimpleNamespace(
            device="cpu", local_rank=-1, distill_ce=False, distill_db=True, distill_de=True, distill_col=False,
            train_de=train_de, temperature=T, layer_temperature=0.5, distill_de_db_layer_score=True,
            distill_col_db_layer_score=False, distill_ce_db_layer_score=False, distill_col_db_attention=False,
            distill_de_db_attention=False, distill_ce_db_attention=False, layer_score_reweight=reweight,
            distill_para_db=1.0, distill_para_de=0.7, distill_para_de_db_dis=1.5, distill_para_de_db_layer_score=0.3)
        batch = {"de_ctx_start_end": torch.zeros(GQ, group + 1, 2, dtype=torch.long)}
        loss, loss_dict = util.distill_loss(args, None, t["q_de"], t["q_db"], None, t["c_de"], t["c_db"], None, sel_db, sel_de,
                                            None, t["qL_de"], t["cL_de"], t["qL_db"], t["cL_db"], None, None, None, None,
                                            None, batch)
        loss.backward()
        res[tag + "_loss"] = np.array([loss.item()])
        res[tag + "_cfg"] = np.array([float(train_de), float(reweight), T, 0.5, 1.0, 0.7, 1.5, 0.3])
        for n, v in t.items():
            res[f"{tag}_d_{n}"] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy().copy()
        print("lead_distill", tag, loss.item(), {k_: round(float(v_), 5) for k_, v_ in loss_dict.items()})
    # ---- cross-encoder teacher variant (--distill_ce): local [1+n] blocks, CE trained too, per-layer CE logits ----
    ce_logits = (2.0 * rng.standard_normal((GQ, group))).astype(np.float32)
    ce_layers = (2.0 * rng.standard_normal((12, GQ, group))).astype(np.float32)
    res.update(ce_logits=ce_logits, ce_layers=ce_layers)
    t = {n: torch.from_numpy(v.copy()).requires_grad_(True) for n, v in
         dict(q_db=arrs["q_db"], c_db=arrs["c_db"], qL_db=arrs["qL_db"], cL_db=arrs["cL_db"], ce_logits=ce_logits,
              ce_layers=ce_layers).items()}
    args = types.SimpleNamespace(
        device="cpu", local_rank=-1, distill_ce=True, distill_db=True, distill_de=False, distill_col=False, train_ce=True,
        temperature=2.0, layer_temperature=0.5, distill_de_db_layer_score=False, distill_col_db_layer_score=False,
        distill_ce_db_layer_score=True, distill_col_db_attention=False, distill_de_db_attention=False,
        distill_ce_db_attention=False, layer_score_reweight=True, distill_para_db=1.0, distill_para_ce=0.6,
        distill_para_ce_db_dis=1.2, distill_para_ce_db_layer_score=0.3)
    batch = {"de_ctx_start_end": torch.zeros(GQ, group + 1, 2, dtype=torch.long)}
    loss, loss_dict = util.distill_loss(args, None, None, t["q_db"], None, None, t["c_db"], None, sel_db, sel_de, None,
                                        None, None, t["qL_db"], t["cL_db"], None, None, t["ce_layers"], None,
                                        t["ce_logits"], batch)
    loss.backward()
    res["ce_loss"] = np.array([loss.item()])
    res["ce_cfg"] = np.array([2.0, 0.5, 1.0, 0.6, 1.2])
    for n, v in t.items():
        res[f"ce_d_{n}"] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy().copy()
    print("lead_distill ce", loss.item(), {k_: round(float(v_), 5) for k_, v_ in loss_dict.items()})
    np.savez_compressed(GOLD / "lead_distill.npz", **res)


def golden_lead_col_distill():
    """the real LEAD distill_loss with a ColBERT teacher (--distill_col --distill_db --distill_col_db_layer_score: the
    shipped distill_from_12cb_to_6de.sh), late-interaction scores from colbert_score on final and per-layer token states"""
    lead_shim()
    util = import_from(REF / "LEAD", "util")
    rng = np.random.RandomState(43)
    GQ, group, H, Sq, Sd = 4, 3, 16, 5, 7
    mk = lambda *shape: (rng.standard_normal(shape) * 0.5).astype(np.float32)
    unit = lambda a: a / np.linalg.norm(a, axis=-1, keepdims=True)
    arrs = dict(q_db=mk(GQ, H), c_db=mk(GQ * group, H), q_col=unit(mk(GQ, Sq, H)), c_col=unit(mk(GQ * group, Sd, H)),
                qL_db=mk(GQ, 6, 1, H), cL_db=mk(GQ * group, 6, 1, H), qL_col=mk(GQ, 12, Sq, H), cL_col=mk(GQ * group, 12, Sd, H))
    lens = rng.randint(2, Sd + 1, size=GQ * group)
    doc_mask = (np.arange(Sd)[None, :] < lens[:, None]).astype(np.int64)
    sel_db, sel_col = [0, 2, 5], [3, 8, 11]
    res = dict(arrs, doc_mask=doc_mask, sel_db=np.array(sel_db), sel_col=np.array(sel_col), group=np.array([group]))
    for tag, train_col, reweight, T in (("both", True, True, 2.0), ("frozen", False, False, 1.0)):
        t = {n: torch.from_numpy(v.copy()).requires_grad_(True) for n, v in arrs.items()}
        args = types.SimpleNamespace(
            device="cpu", local_rank=-1, distill_ce=False, distill_db=True, distill_de=False, distill_col=True,
            train_col=train_col, temperature=T, layer_temperature=0.5, distill_de_db_layer_score=False,
            distill_col_db_layer_score=True, distill_ce_db_layer_score=False, distill_col_db_attention=False,
            distill_de_db_attention=False, distill_ce_db_attention=False, layer_score_reweight=reweight,
            distill_para_db=1.0, distill_para_col=0.8, distill_para_col_db_dis=1.3, distill_para_col_db_layer_score=0.4)
        batch = {"de_ctx_start_end": torch.zeros(GQ, group + 1, 2, dtype=torch.long)}
        loss, loss_dict = util.distill_loss(args, t["q_col"], None, t["q_db"], t["c_col"], None, t["c_db"], None, sel_db,
                                            sel_col, None, None, None, t["qL_db"], t["cL_db"], t["qL_col"], t["cL_col"],
                                            None, torch.from_numpy(doc_mask), None, batch)
        loss.backward()
        res[tag + "_loss"] = np.array([loss.item()])
        res[tag + "_cfg"] = np.array([float(train_col), float(reweight), T, 0.5, 1.0, 0.8, 1.3, 0.4])
        for n, v in t.items():
            res[f"{tag}_d_{n}"] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy().copy()
        print("lead_col_distill", tag, loss.item(), {k_: round(float(v_), 5) for k_, v_ in loss_dict.items()})
    # ---- warm-up loss of the ColBERT tower: get_loss_dual's late-interaction branch (util.py:416-451, fill -1e4) ----
    t = {n: torch.from_numpy(arrs[n].copy()).requires_grad_(True) for n in ("q_col", "c_col")}
    args = types.SimpleNamespace(device="cpu", local_rank=-1, model_type="colbert")
    loss = util.get_loss_dual(args, t["q_col"], t["c_col"], torch.from_numpy(doc_mask))
    loss.backward()
    res["warm_loss"] = np.array([loss.item()])
    res["warm_d_q_col"], res["warm_d_c_col"] = t["q_col"].grad.numpy().copy(), t["c_col"].grad.numpy().copy()
    ce_logits = mk(GQ, group)
    tl = torch.from_numpy(ce_logits.copy()).requires_grad_(True)
    lc = util.get_loss_cross(args, tl)
    lc.backward()
    res.update(ce_logits=ce_logits, ce_warm_loss=np.array([lc.item()]), ce_warm_d=tl.grad.numpy().copy())
    np.savez_compressed(GOLD / "lead_col_distill.npz", **res)


def golden_lead_virt():
    """LEAD/util.py:504-537 query_doc_attention_map + virt_loss, the reference functions themselves, on synthetic per-layer
    hidden states of a teacher and a student: maps, loss and the gradients with respect to all four inputs"""
    lead_shim()
    util = sys.modules.get("util") or import_from(REF / "LEAD", "util")
    rng = np.random.RandomState(47)
    Q, D, Sq, Sd, H = 3, 6, 5, 9, 16
    mk = lambda *shape: (rng.standard_normal(shape) * 0.7).astype(np.float32)
    arrs = dict(q_t=mk(Q, 12, Sq, H), d_t=mk(D, 12, Sd, H), q_s=mk(Q, 6, Sq, H), d_s=mk(D, 6, Sd, H))
    query_len = rng.randint(2, Sq + 1, size=Q).tolist()
    doc_len = rng.randint(3, Sd + 1, size=D).tolist()
    sel_s, sel_t = [1, 4], [3, 10]
    t = {n: torch.from_numpy(v.copy()).requires_grad_(True) for n, v in arrs.items()}
    args = types.SimpleNamespace(device="cpu", temperature=2.0)
    maps_t, _, tgt_t = util.query_doc_attention_map(args, t["q_t"], t["d_t"], query_len, doc_len, sel_t)
    maps_s, lsm_s, _ = util.query_doc_attention_map(args, t["q_s"], t["d_s"], query_len, doc_len, sel_s)
    loss = util.virt_loss(maps_t, maps_s, 0.7)
    loss.backward()
    res = dict(arrs, query_len=np.array(query_len), doc_len=np.array(doc_len), sel_s=np.array(sel_s), sel_t=np.array(sel_t),
               loss=np.array([loss.item()]), map_t0=maps_t[0].detach().numpy(), lsm_s1=lsm_s[1].detach().numpy(),
               tgt_t1=tgt_t[1].detach().numpy())
    for n, v in t.items():
        res["d_" + n] = v.grad.numpy().copy()
    np.savez_compressed(GOLD / "lead_virt.npz", **res)
    print("lead_virt loss", loss.item())


def golden_lead_colbert():
    """LEAD/models.py HFColBertEncoder (the ColBERT teacher tower) on a tiny BERT: normalised token embeddings, every
    layer's FULL hidden state, and the gradients of a loss built from the reference's own colbert_score on the final
    output and on an intermediate layer (what --distill_col_db_layer_score differentiates through)."""
    lead_shim()
    models = import_from(REF / "LEAD", "models")
    util = sys.modules.get("util") or import_from(REF / "LEAD", "util")
    mb = sys.modules["modeling_bert"]
    if not hasattr(mb.BertPreTrainedModel, "get_head_mask"):
        mb.BertPreTrainedModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    cfg = bert_config(num_layers=3, hidden=128, heads=2, intermediate=256, vocab=512, max_pos=64)
    params = synth_params(cfg, seed=51)
    hcfg = hf_bert_config(cfg)
    hcfg.add_linear = False
    hcfg.output_hidden_states = True
    model = models.HFColBertEncoder(hcfg, mask_punctuation=False)
    load_params(model, params)
    model.eval()
    q_ids, q_mask = synth_tokens(2, 8, cfg["vocab"], seed=52, ragged=True, min_len=3)
    d_ids, d_mask = synth_tokens(6, 24, cfg["vocab"], seed=53, ragged=True, min_len=5)
    tq, td = (torch.from_numpy(q_ids), torch.from_numpy(q_mask)), (torch.from_numpy(d_ids), torch.from_numpy(d_mask))
    q_out, q_hid, _ = model("colbert_query", "cpu", input_ids=tq[0], attention_mask=tq[1])
    d_out, d_hid, _ = model("colbert_doc", "cpu", input_ids=td[0], attention_mask=td[1])
    target = torch.tensor([0, 3])
    scores = util.colbert_score(q_out, d_out, td[1], 2)
    layer = 1
    layer_scores = util.colbert_score(q_hid[:, layer], d_hid[:, layer], td[1], 2) / 64.0
    lsm = torch.nn.functional.log_softmax
    loss = torch.nn.functional.nll_loss(lsm(scores, dim=-1), target) + \
        0.5 * torch.nn.functional.nll_loss(lsm(layer_scores, dim=-1), target)
    loss.backward()
    res = dict(q_ids=q_ids, q_mask=q_mask, d_ids=d_ids, d_mask=d_mask, seed=np.array([51]), layer=np.array([layer]),
               q_out=q_out.detach().numpy(), d_out=d_out.detach().numpy(), q_hid=q_hid.detach().numpy(),
               d_hid=d_hid.detach().numpy(), scores=scores.detach().numpy(), layer_scores=layer_scores.detach().numpy(),
               loss=np.array([loss.item()]))
    for n, p_ in model.named_parameters():
        res["grad/" + n] = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).numpy().copy()
    np.savez_compressed(GOLD / "lead_colbert_tiny.npz", **res)
    print("lead_colbert_tiny loss", loss.item(), "hidden", tuple(q_hid.shape), tuple(d_hid.shape))


def main():
    only = sys.argv[1:]
    if only == ["lead_col"]:
        golden_lead_col_distill()
        golden_lead_colbert()
        golden_lead_virt()
        return
    if only == ["lead"]:
        GOLD.mkdir(parents=True, exist_ok=True)
        golden_lead_distill()
        return
    if only == ["select_layer"]:
        golden_lead_select_layer()
        return
    if only == ["data"]:
        golden_batch_layout()
        return
    if only == ["roberta"]:
        golden_roberta(import_from(REF / "SimANS", "model.models"))
        return
    if only == ["cfg1_bf16"]:
        torch.manual_seed(0)
        golden_biencoder_cfg1_bf16_autocast(import_from(REF / "SimANS", "model.models"))
        return
    if only == ["prod_kd"]:
        golden_losses_prod_kd(import_from(REF / "PROD" / "ProD_KD", "model.models"))
        return
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    GOLD.mkdir(parents=True, exist_ok=True)
    simans = import_from(REF / "SimANS", "model.models")
    golden_encoder_tiny(simans)
    golden_reranker_tiny(simans)
    prod = import_from(REF / "PROD" / "ProD_KD", "model.models")
    simans = import_from(REF / "SimANS", "model.models")
    golden_losses(simans, prod)
    golden_losses_prod_kd(prod)
    golden_lead_distilbert()
    if "lead" in only or not only:
        golden_lead_distill()
    simans = import_from(REF / "SimANS", "model.models")
    golden_biencoder_cfg1(simans)
    golden_biencoder_cfg1_bf16_autocast(simans)
    golden_roberta(simans)
    golden_batch_layout()
    golden_lead_select_layer()
    golden_lead_col_distill()
    golden_lead_colbert()
    golden_lead_virt()


if __name__ == "__main__":
    main()
