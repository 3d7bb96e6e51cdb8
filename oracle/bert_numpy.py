"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement, in numpy float64 with a hand-derived backward, of the encoder the reference's BiEncoder wraps:

  * BERT      : `HFBertEncoder.forward` -> transformers `BertModel.forward`
                (/root/reference/SimANS/model/models.py:58-82; the arithmetic is in the un-vendored `transformers`
                package — the reference pins no version (README: "conda install transformers", v4.x era); the
                in-tree restatement is /root/reference/LEAD/modeling_bert.py: embeddings :201-240, self-attention
                :275-374, BertSelfOutput :384-388, BertIntermediate :449-452, BertOutput :462-466, pooler :655-660).
  * DistilBERT: /root/reference/LEAD/modeling_distilbert.py: Embeddings :93-135, MultiHeadSelfAttention :137-234,
                TransformerBlock :260-319 (same post-LN block, no token types, no pooler).

Parity pin: `oracle/make_golden.py` runs the reference's own classes (imported from /root/reference) on seeded
weights/inputs and stores their outputs under tests/golden/; tests/test_oracle_golden.py checks this file against
those fixtures (the reference ships no tests or golden vectors of its own — SURVEY §4, §8c).
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is in the image; fall back to math.erf
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


# ---------------------------------------------------------------------------------------------------------
# configuration + canonical parameter naming (HF state_dict names)
# ---------------------------------------------------------------------------------------------------------
def bert_config(num_layers=12, hidden=768, heads=12, intermediate=3072, vocab=30522, max_pos=512, type_vocab=2,
                ln_eps=1e-12, arch="bert", pad_token_id=0, roberta_positions=False):
    """roberta_positions: HF RobertaModel's position ids (transformers modeling_roberta.create_position_ids_from_input_ids,
    used through /root/reference/SimANS/model/models.py:168-192 HFRobertaEncoder and :333-352 RobertaDot):
    pad + cumsum(ids != pad) for real tokens, pad for padding; the pad row of BOTH embedding tables gets no gradient."""
    return dict(arch=arch, num_layers=num_layers, hidden=hidden, heads=heads, intermediate=intermediate, vocab=vocab,
                max_pos=max_pos, type_vocab=(type_vocab if arch == "bert" else 0), ln_eps=ln_eps,
                pad_token_id=pad_token_id, roberta_positions=roberta_positions)


def roberta_config(num_layers=12, hidden=768, heads=12, intermediate=3072, vocab=50265, max_pos=514, ln_eps=1e-5):
    """RobertaConfig defaults of roberta-base: one token type, pad id 1, positions offset by the pad id"""
    return bert_config(num_layers, hidden, heads, intermediate, vocab, max_pos, type_vocab=1, ln_eps=ln_eps,
                       pad_token_id=1, roberta_positions=True)


def position_ids(cfg, ids):
    ids = np.asarray(ids)
    if not cfg.get("roberta_positions", False):
        return np.broadcast_to(np.arange(ids.shape[1]), ids.shape)
    pad = cfg["pad_token_id"]
    m = (ids != pad).astype(np.int64)
    return np.cumsum(m, axis=1) * m + pad


def param_names(cfg):
    """HF state_dict keys in the canonical order of include/simx_b200.h."""
    L = cfg["num_layers"]
    if cfg["arch"] == "bert":
        names = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
                 "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"]
        for l in range(L):
            p = f"encoder.layer.{l}."
            names += [p + "attention.self.query.weight", p + "attention.self.query.bias",
                      p + "attention.self.key.weight", p + "attention.self.key.bias",
                      p + "attention.self.value.weight", p + "attention.self.value.bias",
                      p + "attention.output.dense.weight", p + "attention.output.dense.bias",
                      p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias",
                      p + "intermediate.dense.weight", p + "intermediate.dense.bias",
                      p + "output.dense.weight", p + "output.dense.bias",
                      p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"]
        names += ["pooler.dense.weight", "pooler.dense.bias"]
    else:
        names = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
                 "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"]
        for l in range(L):
            p = f"transformer.layer.{l}."
            names += [p + "attention.q_lin.weight", p + "attention.q_lin.bias",
                      p + "attention.k_lin.weight", p + "attention.k_lin.bias",
                      p + "attention.v_lin.weight", p + "attention.v_lin.bias",
                      p + "attention.out_lin.weight", p + "attention.out_lin.bias",
                      p + "sa_layer_norm.weight", p + "sa_layer_norm.bias",
                      p + "ffn.lin1.weight", p + "ffn.lin1.bias",
                      p + "ffn.lin2.weight", p + "ffn.lin2.bias",
                      p + "output_layer_norm.weight", p + "output_layer_norm.bias"]
    return names


def param_shapes(cfg):
    H, I, L = cfg["hidden"], cfg["intermediate"], cfg["num_layers"]
    shapes = [(cfg["vocab"], H), (cfg["max_pos"], H)]
    if cfg["arch"] == "bert":
        shapes.append((cfg["type_vocab"], H))
    shapes += [(H,), (H,)]
    for _ in range(L):
        shapes += [(H, H), (H,)] * 4 + [(H,), (H,)] + [(I, H), (I,), (H, I), (H,)] + [(H,), (H,)]
    if cfg["arch"] == "bert":
        shapes += [(H, H), (H,)]
    return shapes


def synth_params(cfg, seed=0, perturb=True):
    """Deterministic synthetic weights (float32).  perturb=False: HF default init (N(0,0.02) weights, zero biases,
    LayerNorm 1/0, SimANS/model/models.py:457-466).  perturb=True additionally randomises biases and LayerNorm
    affine parameters so that every gradient path is exercised by the parity tests."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape in zip(param_names(cfg), param_shapes(cfg)):
        is_ln = "LayerNorm" in name or "layer_norm" in name
        if is_ln and name.endswith("weight"):
            v = np.ones(shape) + (0.1 * rng.standard_normal(shape) if perturb else 0.0)
        elif is_ln or name.endswith("bias"):
            v = 0.02 * rng.standard_normal(shape) if perturb else np.zeros(shape)
        else:
            v = 0.02 * rng.standard_normal(shape)
        out[name] = v.astype(np.float32)
    return out


def synth_tokens(B, S, vocab, seed=1234, ragged=False, min_len=4):
    """ids ~ U{1000..vocab-1}, [CLS]=101 first, [SEP]=102 last valid, zero padding; mask = ids != 0
    (SURVEY §8d synthetic inputs; MARCO_until_new.py:240-260 builds masks the same way)."""
    rng = np.random.RandomState(seed)
    lo = min(1000, max(1, vocab // 2))
    ids = rng.randint(lo, vocab, size=(B, S)).astype(np.int64)
    ids[:, 0] = min(101, vocab - 1)
    lens = rng.randint(min(min_len, S), S + 1, size=B) if ragged else np.full(B, S)
    for b in range(B):
        ids[b, lens[b] - 1] = min(102, vocab - 1)
        ids[b, lens[b]:] = 0
    mask = (ids != 0).astype(np.int64)
    return ids, mask


# ---------------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------------
def gelu(x):
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def gelu_grad(x):
    return 0.5 * (1.0 + _erf(x / math.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def layer_norm(z, g, b, eps):
    mu = z.mean(-1, keepdims=True)
    var = ((z - mu) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (z - mu) * rstd
    return xhat * g + b, (xhat, rstd)


def layer_norm_bwd(dy, g, cache):
    xhat, rstd = cache
    dg = (dy * xhat).reshape(-1, xhat.shape[-1]).sum(0)
    db = dy.reshape(-1, xhat.shape[-1]).sum(0)
    dxh = dy * g
    dz = rstd * (dxh - dxh.mean(-1, keepdims=True) - xhat * (dxh * xhat).mean(-1, keepdims=True))
    return dz, dg, db


class EncoderOracle:
    """forward()/backward() of one tower.  `params`: dict HF-name -> array."""

    def __init__(self, cfg, params, dtype=np.float64):
        self.cfg = cfg
        self.dtype = dtype
        self.names = param_names(cfg)
        self.p = [np.asarray(params[n], dtype=dtype) for n in self.names]
        self.base = 5 if cfg["arch"] == "bert" else 4

    def _layer_params(self, l):
        return self.p[self.base + 16 * l: self.base + 16 * (l + 1)]

    # ---- forward ---------------------------------------------------------------------------------------
    def forward(self, ids, mask=None, dropout=None, inputs_embeds=None):
        """dropout: None (eval / parity configuration) or dict(p_hidden=, p_attn=, seed=): the masks of
        oracle/dropout_numpy.py at the HF sites (LEAD/modeling_bert.py:239, :358, :386, :464; DistilBERT has no dropout
        after the attention output projection, LEAD/modeling_distilbert.py:254-319)."""
        cfg = self.cfg
        if inputs_embeds is not None:
            # HF BertEmbeddings.forward(inputs_embeds=...): the caller's [B,S,H] rows replace word_embeddings(input_ids)
            # (LEAD/modeling_bert.py:222-227); `ids` is only used for its shape
            inputs_embeds = np.asarray(inputs_embeds, dtype=self.dtype)
            ids = np.zeros(inputs_embeds.shape[:2], dtype=np.int64)
        B, S = ids.shape
        H, nh = cfg["hidden"], cfg["heads"]
        d = H // nh
        from oracle import dropout_numpy as dn
        dr = dropout or dict(p_hidden=0.0, p_attn=0.0, seed=0)
        ph, pa, seed = float(dr.get("p_hidden", 0.0)), float(dr.get("p_attn", 0.0)), int(dr.get("seed", 0))
        hmask = lambda site, layer: dn.hidden_mask(seed, site, layer, B * S, H, ph).reshape(B, S, H)
        if mask is None:
            mask = np.ones_like(ids)
        word, pos = self.p[0], self.p[1]
        pos_ids = position_ids(cfg, ids)
        emb = (word[ids] if inputs_embeds is None else inputs_embeds) + pos[pos_ids]
        if cfg["arch"] == "bert":
            emb = emb + self.p[2][0][None, None]
            g, b = self.p[3], self.p[4]
        else:
            g, b = self.p[2], self.p[3]
        x, ln0 = layer_norm(emb, g, b, cfg["ln_eps"])
        m0 = hmask(dn.SITE_EMBED, 0)
        x = x * m0
        self.cache = dict(ids=ids, mask=mask, ln0=ln0, layers=[], pos_ids=pos_ids, m0=m0, embeds=inputs_embeds is not None)
        hidden = [x]
        neg = np.where(mask[:, None, None, :] > 0, 0.0, -np.inf)
        for l in range(cfg["num_layers"]):
            (wq, bq, wk, bk, wv, bv, wo, bo, g1, b1, wi, bi, w2, b2, g2, bb2) = self._layer_params(l)
            q = (x @ wq.T + bq).reshape(B, S, nh, d).transpose(0, 2, 1, 3)
            k = (x @ wk.T + bk).reshape(B, S, nh, d).transpose(0, 2, 1, 3)
            v = (x @ wv.T + bv).reshape(B, S, nh, d).transpose(0, 2, 1, 3)
            s = q @ k.transpose(0, 1, 3, 2) / math.sqrt(d) + neg
            s = s - s.max(-1, keepdims=True)
            e = np.exp(s)
            p = e / e.sum(-1, keepdims=True)
            ma = dn.attn_mask(seed, l, B, nh, S, pa)
            m2 = hmask(dn.SITE_ATTN_OUT, l) if cfg["arch"] == "bert" else np.ones((B, S, H))
            m3 = hmask(dn.SITE_FFN_OUT, l)
            ctx = ((p * ma) @ v).transpose(0, 2, 1, 3).reshape(B, S, H)
            z1 = (ctx @ wo.T + bo) * m2 + x
            x1, ln1 = layer_norm(z1, g1, b1, cfg["ln_eps"])
            h = x1 @ wi.T + bi
            a = gelu(h)
            z2 = (a @ w2.T + b2) * m3 + x1
            x2, ln2 = layer_norm(z2, g2, bb2, cfg["ln_eps"])
            self.cache["layers"].append(dict(x=x, q=q, k=k, v=v, p=p, ctx=ctx, ln1=ln1, x1=x1, h=h, a=a, ln2=ln2,
                                             ma=ma, m2=m2, m3=m3))
            x = x2
            hidden.append(x)
        self.hidden = hidden
        cls = np.stack([hh[:, 0, :] for hh in hidden], axis=1)  # [B, L+1, H]
        return dict(cls=cls, seq=x, hidden=hidden)

    # ---- backward --------------------------------------------------------------------------------------
    def backward(self, d_cls=None, d_seq=None, d_hidden=None):
        """d_cls [B, L+1, H] cotangent of forward()['cls']; d_seq [B,S,H] cotangent of the last hidden state; d_hidden
        [B, L+1, S, H] cotangent of forward()['hidden'] stacked on axis 1 (HF output_hidden_states: what LEAD's ColBERT
        layer scores and attention maps differentiate through, LEAD/util.py:504-529, :657-660).
        Returns dict HF-name -> gradient (dense for every parameter; pooler gradients are zeros)."""
        cfg = self.cfg
        c = self.cache
        ids = c["ids"]
        B, S = ids.shape
        H, nh, L = cfg["hidden"], cfg["heads"], cfg["num_layers"]
        d = H // nh
        grads = [np.zeros_like(w) for w in self.p]
        dx = np.zeros((B, S, H), dtype=self.dtype)
        if d_seq is not None:
            dx += d_seq
        if d_cls is not None:
            dx[:, 0, :] += d_cls[:, L, :]
        if d_hidden is not None:
            dx += d_hidden[:, L]
        for l in range(L - 1, -1, -1):
            (wq, bq, wk, bk, wv, bv, wo, bo, g1, b1, wi, bi, w2, b2, g2, bb2) = self._layer_params(l)
            lc = c["layers"][l]
            gi = self.base + 16 * l
            dz2, grads[gi + 14], grads[gi + 15] = layer_norm_bwd(dx, g2, lc["ln2"])
            dz2m = dz2 * lc["m3"]  # into the dense output (dropout sits between the dense and the residual add)
            f2 = dz2m.reshape(-1, H)
            grads[gi + 13] = f2.sum(0)
            grads[gi + 12] = f2.T @ lc["a"].reshape(-1, lc["a"].shape[-1])
            dh = (dz2m @ w2) * gelu_grad(lc["h"])
            fh = dh.reshape(-1, dh.shape[-1])
            grads[gi + 11] = fh.sum(0)
            grads[gi + 10] = fh.T @ lc["x1"].reshape(-1, H)
            dx1 = dh @ wi + dz2
            dz1, grads[gi + 8], grads[gi + 9] = layer_norm_bwd(dx1, g1, lc["ln1"])
            dz1m = dz1 * lc["m2"]
            f1 = dz1m.reshape(-1, H)
            grads[gi + 7] = f1.sum(0)
            grads[gi + 6] = f1.T @ lc["ctx"].reshape(-1, H)
            dctx = (dz1m @ wo).reshape(B, S, nh, d).transpose(0, 2, 1, 3)
            p, q, k, v = lc["p"], lc["q"], lc["k"], lc["v"]
            dv = (p * lc["ma"]).transpose(0, 1, 3, 2) @ dctx
            dp = (dctx @ v.transpose(0, 1, 3, 2)) * lc["ma"]
            ds = p * (dp - (dp * p).sum(-1, keepdims=True)) / math.sqrt(d)
            dq = ds @ k
            dk = ds.transpose(0, 1, 3, 2) @ q
            x = lc["x"].reshape(-1, H)
            dxl = dz1.copy()
            for j, (w, dd) in enumerate(((wq, dq), (wk, dk), (wv, dv))):
                f = dd.transpose(0, 2, 1, 3).reshape(-1, H)
                grads[gi + 2 * j] = f.T @ x
                grads[gi + 2 * j + 1] = f.sum(0)
                dxl += (f @ w).reshape(B, S, H)
            dx = dxl
            if d_cls is not None:
                dx[:, 0, :] += d_cls[:, l, :]
            if d_hidden is not None:
                dx = dx + d_hidden[:, l]
        dx = dx * c["m0"]
        if cfg["arch"] == "bert":
            g = self.p[3]
            demb, grads[3], grads[4] = layer_norm_bwd(dx, g, c["ln0"])
            grads[2][0] = demb.reshape(-1, H).sum(0)
        else:
            g = self.p[2]
            demb, grads[2], grads[3] = layer_norm_bwd(dx, g, c["ln0"])
        pad = cfg.get("pad_token_id", 0)
        if cfg.get("roberta_positions", False):
            np.add.at(grads[1], c["pos_ids"].reshape(-1), demb.reshape(-1, H))
            grads[1][pad] = 0.0  # RobertaEmbeddings: position_embeddings has padding_idx too
        else:
            grads[1][:S] = demb.sum(0)
        self.d_inputs_embeds = demb if c["embeds"] else None
        if not c["embeds"]:
            np.add.at(grads[0], ids.reshape(-1), demb.reshape(-1, H))
        grads[0][pad] = 0.0  # nn.Embedding(padding_idx=pad): the pad row never receives gradient
        return {n: g_ for n, g_ in zip(self.names, grads)}
