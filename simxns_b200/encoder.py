"""Host side of one encoder tower: HF-named fp32 master parameters (`state_dict` identical to
transformers' BertModel / DistilBertModel) + the autograd.Function that runs the tower on
libsimx_b200 (tcgen05 GEMMs, attention, LayerNorm kernels).  PyTorch is plumbing here: it owns device memory,
streams, the autograd graph root and the parameters; every FLOP of the tower happens inside the C ABI calls.

Replaces: HFBertEncoder over transformers.BertModel (/root/reference/SimANS/model/models.py:58-82,
/root/reference/PROD/ProD_KD/model/models.py:13-82, /root/reference/LEAD/models.py:31-196).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import _lib


@dataclass
class EncoderConfig:
    arch: str = "bert"  # "bert" | "distilbert"
    num_layers: int = 12
    hidden: int = 768
    heads: int = 12
    intermediate: int = 3072
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12
    init_range: float = 0.02
    pad_token_id: int = 0            # row of the embedding tables that never receives gradient (nn.Embedding padding_idx)
    roberta_positions: bool = False  # HF RobertaModel position ids: pad + cumsum(ids != pad), pad for padding
    pooler: bool = True              # False: RobertaModel(add_pooling_layer=False) — no `pooler.*` parameters exposed

    def c_struct(self, precision: int, hidden_dropout: float = 0.0, attn_dropout: float = 0.0, dropout_seed: int = 0,
                 save_policy: int = 0, id_bytes: int = 8) -> _lib.EncoderCfg:
        return _lib.EncoderCfg(
            arch=_lib.ARCH_BERT if self.arch == "bert" else _lib.ARCH_DISTILBERT,
            num_layers=self.num_layers, hidden=self.hidden, heads=self.heads, intermediate=self.intermediate,
            vocab=self.vocab, max_pos=self.max_pos, type_vocab=self.type_vocab if self.arch == "bert" else 0,
            precision=precision, ln_eps=self.ln_eps, hidden_dropout=hidden_dropout, attn_dropout=attn_dropout,
            dropout_seed=dropout_seed, save_policy=save_policy, id_bytes=id_bytes)

    @staticmethod
    def from_hf(cfg) -> "EncoderConfig":
        """from a transformers BertConfig / DistilBertConfig (or any object with the same attributes)"""
        if getattr(cfg, "model_type", "bert") == "distilbert" or hasattr(cfg, "n_layers"):
            return EncoderConfig(arch="distilbert", num_layers=cfg.n_layers, hidden=cfg.dim, heads=cfg.n_heads,
                                 intermediate=cfg.hidden_dim, vocab=cfg.vocab_size, max_pos=cfg.max_position_embeddings,
                                 type_vocab=0, ln_eps=1e-12, init_range=getattr(cfg, "initializer_range", 0.02))
        roberta = getattr(cfg, "model_type", "bert") in ("roberta", "xlm-roberta")
        act = getattr(cfg, "hidden_act", "gelu")
        if act not in ("gelu",):
            raise NotImplementedError(f"hidden_act={act!r}: only the exact erf GELU of BERT is implemented")
        if getattr(cfg, "position_embedding_type", "absolute") != "absolute":
            raise NotImplementedError("only absolute position embeddings are implemented")
        return EncoderConfig(arch="bert", num_layers=cfg.num_hidden_layers, hidden=cfg.hidden_size,
                             heads=cfg.num_attention_heads, intermediate=cfg.intermediate_size, vocab=cfg.vocab_size,
                             max_pos=cfg.max_position_embeddings, type_vocab=cfg.type_vocab_size,
                             ln_eps=cfg.layer_norm_eps, init_range=getattr(cfg, "initializer_range", 0.02),
                             pad_token_id=(cfg.pad_token_id if roberta else 0), roberta_positions=roberta)


# Every torch.optim.Optimizer.step() — including subclasses that write through `p.data` and therefore never bump
# `p._version`, like the transformers.AdamW of the reference drivers — advances this counter; the bf16 operand cache of
# each tower is keyed on it (and on the parameters' data pointers / versions).
_OPTIMIZER_STEPS = [0]


def _count_optimizer_step(optimizer, args, kwargs):
    _OPTIMIZER_STEPS[0] += 1


from torch.optim.optimizer import register_optimizer_step_post_hook  # noqa: E402

register_optimizer_step_post_hook(_count_optimizer_step)


# Dropout seeds: one 64-bit seed per tower call, derived from torch's seed (the drivers' `set_seed(args)` ->
# torch.manual_seed), the rank, and a per-process call counter; masks are a counter-based function of the seed
# (csrc/dropout.cuh), so backward — and the recomputation under gradient checkpointing — regenerates them.
_DROPOUT_CALLS = [0]


def _mix64(x: int) -> int:
    x &= 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def next_dropout_seed() -> int:
    _DROPOUT_CALLS[0] += 1
    rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank()
    return _mix64(torch.initial_seed() * 0x9E3779B97F4A7C15 + _DROPOUT_CALLS[0] * 0xD1B54A32D192ED03 + (rank << 48) + 1)


def default_precision() -> int:
    """SIMX_PRECISION=bf16 (default, throughput) | bf16x3 / fp32 (3-plane split-bf16, fp32-class parity mode)"""
    v = os.environ.get("SIMX_PRECISION", "bf16").lower()
    if v in ("bf16", "1"):
        return _lib.PREC_BF16
    if v in ("bf16x3", "fp32", "3"):
        return _lib.PREC_BF16X3
    raise ValueError(f"SIMX_PRECISION={v!r}: expected bf16 or bf16x3")


# ---------------------------------------------------------------------------------------------------------
# HF module skeleton: real nn.Embedding / nn.Linear / nn.LayerNorm objects arranged exactly like transformers'
# BertModel / DistilBertModel so that named_parameters(), state_dict() and the optimizer's "bias / LayerNorm.weight"
# no-decay grouping (co_training_marco_train.py:57-69) see the reference's names.  Their forward() is never used.
# ---------------------------------------------------------------------------------------------------------
def _container(**children) -> nn.Module:
    m = nn.Module()
    for k, v in children.items():
        m.add_module(k, v)
    return m


def _build_bert_skeleton(owner: nn.Module, c: EncoderConfig) -> List[nn.Parameter]:
    H, I = c.hidden, c.intermediate
    emb = _container(word_embeddings=nn.Embedding(c.vocab, H, padding_idx=c.pad_token_id),
                     position_embeddings=nn.Embedding(c.max_pos, H,
                                                      padding_idx=c.pad_token_id if c.roberta_positions else None),
                     token_type_embeddings=nn.Embedding(c.type_vocab, H),
                     LayerNorm=nn.LayerNorm(H, eps=c.ln_eps))
    layers = nn.ModuleList()
    order: List[nn.Parameter] = [emb.word_embeddings.weight, emb.position_embeddings.weight,
                                 emb.token_type_embeddings.weight, emb.LayerNorm.weight, emb.LayerNorm.bias]
    for _ in range(c.num_layers):
        att_self = _container(query=nn.Linear(H, H), key=nn.Linear(H, H), value=nn.Linear(H, H))
        att_out = _container(dense=nn.Linear(H, H), LayerNorm=nn.LayerNorm(H, eps=c.ln_eps))
        layer = _container(attention=_container(self=att_self, output=att_out),
                           intermediate=_container(dense=nn.Linear(H, I)),
                           output=_container(dense=nn.Linear(I, H), LayerNorm=nn.LayerNorm(H, eps=c.ln_eps)))
        layers.append(layer)
        order += [att_self.query.weight, att_self.query.bias, att_self.key.weight, att_self.key.bias,
                  att_self.value.weight, att_self.value.bias, att_out.dense.weight, att_out.dense.bias,
                  att_out.LayerNorm.weight, att_out.LayerNorm.bias, layer.intermediate.dense.weight,
                  layer.intermediate.dense.bias, layer.output.dense.weight, layer.output.dense.bias,
                  layer.output.LayerNorm.weight, layer.output.LayerNorm.bias]
    pooler = _container(dense=nn.Linear(H, H))
    order += [pooler.dense.weight, pooler.dense.bias]
    owner.add_module("embeddings", emb)
    owner.add_module("encoder", _container(layer=layers))
    if c.pooler:
        owner.add_module("pooler", pooler)
    else:
        # the kernel's parameter table always has the two (dead) pooler slots: keep zero placeholders that follow the
        # module between devices but are neither parameters nor part of the state_dict
        owner.register_buffer("_no_pooler_w", torch.zeros(H, H), persistent=False)
        owner.register_buffer("_no_pooler_b", torch.zeros(H), persistent=False)
        order[-2:] = [owner._no_pooler_w, owner._no_pooler_b]
    return order


def _build_distilbert_skeleton(owner: nn.Module, c: EncoderConfig) -> List[nn.Parameter]:
    H, I = c.hidden, c.intermediate
    emb = _container(word_embeddings=nn.Embedding(c.vocab, H, padding_idx=0),
                     position_embeddings=nn.Embedding(c.max_pos, H), LayerNorm=nn.LayerNorm(H, eps=c.ln_eps))
    order: List[nn.Parameter] = [emb.word_embeddings.weight, emb.position_embeddings.weight, emb.LayerNorm.weight,
                                 emb.LayerNorm.bias]
    layers = nn.ModuleList()
    for _ in range(c.num_layers):
        att = _container(q_lin=nn.Linear(H, H), k_lin=nn.Linear(H, H), v_lin=nn.Linear(H, H), out_lin=nn.Linear(H, H))
        layer = _container(attention=att, sa_layer_norm=nn.LayerNorm(H, eps=c.ln_eps),
                           ffn=_container(lin1=nn.Linear(H, I), lin2=nn.Linear(I, H)),
                           output_layer_norm=nn.LayerNorm(H, eps=c.ln_eps))
        layers.append(layer)
        order += [att.q_lin.weight, att.q_lin.bias, att.k_lin.weight, att.k_lin.bias, att.v_lin.weight, att.v_lin.bias,
                  att.out_lin.weight, att.out_lin.bias, layer.sa_layer_norm.weight, layer.sa_layer_norm.bias,
                  layer.ffn.lin1.weight, layer.ffn.lin1.bias, layer.ffn.lin2.weight, layer.ffn.lin2.bias,
                  layer.output_layer_norm.weight, layer.output_layer_norm.bias]
    owner.add_module("embeddings", emb)
    owner.add_module("transformer", _container(layer=layers))
    return order


def _hf_init_(owner: nn.Module, init_range: float) -> None:
    """transformers' `_init_weights`: N(0, initializer_range) for Linear/Embedding weights, zero biases and pad row,
    LayerNorm 1/0 (LEAD/modeling_bert.py:745-760)."""
    for m in owner.modules():
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(mean=0.0, std=init_range)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=init_range)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)


class B200EncoderBase(nn.Module):
    """One tower.  Subclasses (models.HFBertEncoder, ...) add the reference's constructor / forward signature."""

    def __init__(self, enc_cfg: EncoderConfig, precision: Optional[int] = None):
        super().__init__()
        self.enc_cfg = enc_cfg
        self.precision = default_precision() if precision is None else precision
        if enc_cfg.arch == "bert":
            self._ordered = _build_bert_skeleton(self, enc_cfg)
        else:
            self._ordered = _build_distilbert_skeleton(self, enc_cfg)
        _hf_init_(self, enc_cfg.init_range)
        self._wbuf: Optional[torch.Tensor] = None
        self._wkey = None
        self._grad_slices = None
        # training-time behaviour of the reference's config (SimANS/model/models.py:66-74): dropout probabilities (applied
        # only while self.training) and --gradient_checkpointing (keep layer inputs only, recompute inside backward)
        self.hidden_dropout_prob = 0.0
        self.attention_probs_dropout_prob = 0.0
        self.gradient_checkpointing = False
        self.forced_dropout_seed: Optional[int] = None  # tests: fixed masks (compared with oracle/dropout_numpy.py)
        self.last_dropout_seed: Optional[int] = None
        self.grad_delivery = os.environ.get("SIMX_GRAD_SLICES", "tower")  # "tower" | "layer" (see encode())
        self._grad_sink = None  # gradsync.FlatGradSync: receives (tower, pass) instead of autograd receiving gradients

    # ---- parameter plumbing ---------------------------------------------------------------------------
    def ordered_parameters(self) -> List[nn.Parameter]:
        """parameters in the canonical order of include/simx_b200.h"""
        return self._ordered

    def _apply(self, fn, *a, **k):  # .to() / .cuda() replace parameter storage: drop cached device state
        self._wbuf, self._wkey = None, None
        out = super()._apply(fn, *a, **k)
        if not self.enc_cfg.pooler and self.enc_cfg.arch == "bert":  # buffers are re-created by _apply: re-link the table
            self._ordered[-2:] = [self._no_pooler_w, self._no_pooler_b]
        return out

    def _c_cfg(self, dropout_seed: int = 0, id_bytes: int = 8, train: Optional[bool] = None) -> _lib.EncoderCfg:
        train = self.training if train is None else train
        ph = float(self.hidden_dropout_prob) if train else 0.0
        pa = float(self.attention_probs_dropout_prob) if train else 0.0
        return self.enc_cfg.c_struct(self.precision, hidden_dropout=ph, attn_dropout=pa, dropout_seed=dropout_seed,
                                     save_policy=_lib.SAVE_LAYER_INPUTS if self.gradient_checkpointing else _lib.SAVE_ALL,
                                     id_bytes=id_bytes)

    def _param_table(self, params) -> "C.Array":
        arr = (C.c_void_p * len(params))()
        for i, p in enumerate(params):
            arr[i] = p.data_ptr()
        return arr

    def _check_params(self, params) -> None:
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("simxns_b200 encoder parameters are on %s: the B200 kernels need CUDA tensors "
                               "(there is no CPU fallback)" % dev)
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise RuntimeError("encoder parameters must be contiguous fp32 tensors on one CUDA device")

    def _weights(self, params) -> torch.Tensor:
        """bf16 / split-bf16 GEMM operands of the fp32 master weights, rebuilt when a master weight may have changed:
        a parameter's (data_ptr, _version) moved, or ANY optimizer stepped since the last build (optimizers that update
        through `p.data` — the transformers.AdamW of the reference drivers does, `p.data.addcdiv_` — never bump
        `_version`; the global step hook above sees them).  An out-of-band `p.data` edit outside an optimizer needs
        `invalidate_weights()`."""
        key = (self.precision, _OPTIMIZER_STEPS[0]) + tuple((p.data_ptr(), p._version) for p in params)
        if self._wbuf is None or key != self._wkey:
            lib = _lib.lib()
            cfg = self._c_cfg(train=False)
            nbytes = lib.simx_encoder_weight_bytes(C.byref(cfg))
            if self._wbuf is None or self._wbuf.numel() != nbytes or self._wbuf.device != params[0].device:
                self._wbuf = torch.empty(nbytes, dtype=torch.uint8, device=params[0].device)
            stream = torch.cuda.current_stream(params[0].device).cuda_stream
            _lib.check(lib.simx_encoder_prepare_weights(C.byref(cfg), self._param_table(params), self._wbuf.data_ptr(),
                                                        stream), "simx_encoder_prepare_weights")
            self._wkey = key
        return self._wbuf

    def invalidate_weights(self) -> None:
        self._wkey = None

    def grad_slices(self) -> List[Tuple[int, int]]:
        if self._grad_slices is None:
            lib = _lib.lib()
            cfg = self._c_cfg(train=False)
            n = lib.simx_encoder_num_params(C.byref(cfg))
            assert n == len(self._ordered), (n, len(self._ordered))
            self._grad_slices = [(lib.simx_encoder_grad_offset(C.byref(cfg), i),
                                  lib.simx_encoder_param_numel(C.byref(cfg), i)) for i in range(n)]
        return self._grad_slices

    # ---- compute ------------------------------------------------------------------------------------------
    def param_groups(self) -> List[List[nn.Parameter]]:
        """[embedding params, layer 0 params, ..., layer L-1 params, pooler params ([] without one)] in the canonical order:
        the granularity at which backward hands gradients to autograd (and thereby to DDP's reducer)"""
        base = 5 if self.enc_cfg.arch == "bert" else 4
        L = self.enc_cfg.num_layers
        o = self._ordered
        return [o[:base]] + [o[base + 16 * l: base + 16 * (l + 1)] for l in range(L)] + [o[base + 16 * L:]]

    def encode(self, input_ids: Optional[torch.Tensor], attention_mask: Optional[torch.Tensor] = None,
               want_sequence: bool = False, inputs_embeds: Optional[torch.Tensor] = None, want_hidden: bool = False):
        """-> (cls_layers [B, L+1, H] fp32, sequence_output [B,S,H] fp32 or None[, hidden [B, L+1, S, H] fp32]).
        cls_layers[:, -1] is the reference's pooled_output; cls_layers[:, l+1] the [CLS] row of layer l.
        want_hidden: a third output, every depth's full hidden state (index 0 = embedding output, l+1 = layer l) with
        gradients — HF `output_hidden_states`, what LEAD's towers return as `all_layer_hidden` (LEAD/models.py:125-138).
        inputs_embeds [B,S,H] (instead of input_ids): HF's `inputs_embeds=` — the word-embedding rows supplied by the caller,
        differentiable (the forward_adv_* paths of BiBertEncoder)."""
        params = self._ordered
        if inputs_embeds is not None:
            if input_ids is not None:
                raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")  # HF's message
            if inputs_embeds.dim() != 3 or inputs_embeds.shape[-1] != self.enc_cfg.hidden:
                raise ValueError(f"inputs_embeds must be [B, S, {self.enc_cfg.hidden}]")
            if self.enc_cfg.roberta_positions:
                raise NotImplementedError("inputs_embeds with RoBERTa position ids is not implemented")
            if want_hidden:
                raise NotImplementedError("want_hidden with inputs_embeds is not implemented")
            needs_grad = torch.is_grad_enabled() and (inputs_embeds.requires_grad or any(p.requires_grad for p in params))
            st = _TowerPass(self, None, attention_mask, bool(want_sequence), needs_grad, inputs_embeds=inputs_embeds)
            if not needs_grad:
                return st.cls, (st.seq if want_sequence else None)
            st.monolithic = True
            cls, seq, _ = _TopNode.apply(st, inputs_embeds, *params)
            return cls, (seq if want_sequence else None)
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, S]")
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        st = _TowerPass(self, input_ids, attention_mask, bool(want_sequence), needs_grad, want_hidden=bool(want_hidden))
        if not needs_grad:
            return (st.cls, (st.seq if want_sequence else None)) + ((st.hidden,) if want_hidden else ())
        # autograd graph.  "tower" (default): one node per tower pass returns all gradients when the pass is complete.
        # "layer": embeddings -> layer 0 -> ... -> layer L-1 -> outputs chained through a dummy token, each node running ONE
        # slice of the backward pass (simx_encoder_backward_layers) and returning that slice's parameter gradients, so a
        # reducer sees them layer by layer.  (Under stock DDP that is SLOWER here: NCCL's kernels cannot share an SM with
        # the persistent one-CTA-per-SM GEMMs, so every bucket all-reduce launched mid-backward stalls a GEMM wave —
        # 2 GPUs: 170.7 vs 159.1 ms per step.  gradsync.FlatGradSync reduces each tower's flat gradient buffer in one
        # call after its pass instead.)
        groups = self.param_groups()
        if self.grad_delivery == "layer":
            token = _EmbedNode.apply(st, *groups[0])
            for l in range(self.enc_cfg.num_layers):
                token = _LayerNode.apply(st, l, token, *groups[1 + l])
            cls, seq, hidden = _TopNode.apply(st, token, *groups[-1])
        else:
            st.monolithic = True
            cls, seq, hidden = _TopNode.apply(st, None, *params)
        return (cls, (seq if want_sequence else None)) + ((hidden,) if want_hidden else ())


class _TowerPass:
    """State of one forward/backward pass of one tower: runs the forward at construction, owns the saved activations, and
    runs the backward slice by slice on behalf of the autograd nodes below."""

    def __init__(self, module: B200EncoderBase, ids, mask, want_seq: bool, needs_grad: bool, inputs_embeds=None,
                 want_hidden: bool = False):
        lib = _lib.lib()
        params = module._ordered
        module._check_params(params)
        dev = params[0].device
        src = ids if inputs_embeds is None else inputs_embeds
        if src.device != dev:
            raise RuntimeError(f"inputs on {src.device}, parameters on {dev}")
        ec = module.enc_cfg
        with torch.no_grad(), torch.cuda.device(dev):
            _lib.require_device()
            embeds = None
            if inputs_embeds is not None:
                embeds = inputs_embeds.detach().to(torch.float32).contiguous()
                idt = torch.int64 if mask is None or mask.dtype != torch.int32 else torch.int32
                B, S = embeds.shape[:2]
            else:
                # int32 batches stay int32 (half the host->device bytes); anything else is the reference's int64
                idt = torch.int32 if ids.dtype == torch.int32 else torch.int64
                ids = ids.to(idt).contiguous()
                B, S = ids.shape
            if mask is not None:
                mask = mask.to(device=dev, dtype=idt).contiguous()
                if tuple(mask.shape) != (B, S):
                    raise ValueError("attention_mask must have the shape of input_ids")
            if S > min(ec.max_pos, 512):
                raise ValueError(f"sequence length {S} exceeds min(max_position_embeddings={ec.max_pos}, 512)")
            pos_ids = None
            if ec.roberta_positions:
                # transformers create_position_ids_from_input_ids: pad + running count of real tokens, pad for padding
                if S + ec.pad_token_id >= ec.max_pos:
                    raise ValueError(f"sequence length {S} needs position {S + ec.pad_token_id} >= "
                                     f"max_position_embeddings={ec.max_pos}")
                real = (ids != ec.pad_token_id).to(idt)
                pos_ids = (torch.cumsum(real, dim=1) * real + ec.pad_token_id).to(idt).contiguous()
            dropping = module.training and (module.hidden_dropout_prob > 0 or module.attention_probs_dropout_prob > 0)
            seed = 0
            if dropping:
                seed = module.forced_dropout_seed if module.forced_dropout_seed is not None else next_dropout_seed()
                module.last_dropout_seed = seed
            cfg = module._c_cfg(dropout_seed=seed, id_bytes=4 if idt == torch.int32 else 8)
            if want_hidden and not needs_grad:
                cfg.save_policy = _lib.SAVE_LAYER_INPUTS  # the hidden states are the layer inputs: keep those and nothing else
            wbuf = module._weights(params)
            table = module._param_table(params)
            stream = torch.cuda.current_stream(dev).cuda_stream
            cls = torch.empty(B, ec.num_layers + 1, ec.hidden, dtype=torch.float32, device=dev)
            seq = torch.empty(B, S, ec.hidden, dtype=torch.float32, device=dev) if want_seq else None
            saved, ws = None, None
            if needs_grad or want_hidden:
                saved = torch.empty(lib.simx_encoder_saved_bytes(C.byref(cfg), B, S), dtype=torch.uint8, device=dev)
            if not needs_grad or module.gradient_checkpointing:
                ws = torch.empty(lib.simx_encoder_workspace_bytes(C.byref(cfg), B, S), dtype=torch.uint8, device=dev)
            if embeds is not None:
                _lib.check(lib.simx_encoder_forward_embeds(C.byref(cfg), table, wbuf.data_ptr(), embeds.data_ptr(),
                                                           _lib.ptr(mask), None, B, S, _lib.ptr(saved), _lib.ptr(ws),
                                                           cls.data_ptr(), _lib.ptr(seq), stream),
                           "simx_encoder_forward_embeds")
            elif pos_ids is None:
                _lib.check(lib.simx_encoder_forward(C.byref(cfg), table, wbuf.data_ptr(), ids.data_ptr(), _lib.ptr(mask), B, S,
                                                    _lib.ptr(saved), _lib.ptr(ws), cls.data_ptr(), _lib.ptr(seq), stream),
                           "simx_encoder_forward")
            else:
                _lib.check(lib.simx_encoder_forward_pos(C.byref(cfg), table, wbuf.data_ptr(), ids.data_ptr(), _lib.ptr(mask),
                                                        pos_ids.data_ptr(), B, S, _lib.ptr(saved), _lib.ptr(ws),
                                                        cls.data_ptr(), _lib.ptr(seq), stream), "simx_encoder_forward_pos")
            hidden = None
            if want_hidden:  # [B, L+1, S, H]: depth l of every sequence is one strided slab
                Lh = ec.num_layers + 1
                hidden = torch.empty(B, Lh, S, ec.hidden, dtype=torch.float32, device=dev)
                for l in range(Lh):
                    _lib.check(lib.simx_encoder_hidden_state(C.byref(cfg), saved.data_ptr(), B, S, l, hidden[:, l].data_ptr(),
                                                             Lh * S * ec.hidden, stream), "simx_encoder_hidden_state")
                if not needs_grad:
                    saved = None
        self.hidden, self.d_hidden, self.has_hidden = hidden, None, hidden is not None
        self.module, self.cfg, self.dev = module, cfg, dev
        self.ids, self.mask, self.pos_ids, self.shape = ids, mask, pos_ids, (B, S)
        self.embeds = embeds is not None  # backward also produces d(loss)/d(inputs_embeds)
        self.d_embeds = None
        self.cls, self.seq, self.want_seq = cls, seq, want_seq
        self.saved, self.wbuf, self.params = saved, wbuf, params
        self.ws = None      # backward workspace: holds the running cotangent between slices
        self.flat = None    # flat fp32 gradient buffer of this pass
        self.d_cls = self.d_seq = None
        self.next_hi = ec.num_layers  # the slice that must run next is [.., next_hi)
        self.monolithic = False
        self.sink = module._grad_sink if needs_grad else None

    def token(self) -> torch.Tensor:
        """the empty tensor that chains the autograd nodes of one pass"""
        return torch.empty(0, dtype=torch.float32, device=self.dev)

    def begin_backward(self, d_cls, d_seq, d_hidden=None) -> bool:
        """called by the top node; False when no cotangent reaches the tower"""
        if self.saved is None:
            raise RuntimeError("encoder backward without saved activations (already freed by an earlier backward, or the "
                               "forward ran under no_grad)")
        if d_hidden is not None and not self.has_hidden:
            d_hidden = None
        if d_cls is None and (d_seq is None or not self.want_seq) and d_hidden is None:
            return False
        lib = _lib.lib()
        B, S = self.shape
        with torch.cuda.device(self.dev):
            self.d_cls = None if d_cls is None else d_cls.to(torch.float32).contiguous()
            self.d_seq = d_seq.to(torch.float32).contiguous() if (d_seq is not None and self.want_seq) else None
            if d_hidden is not None:
                # the cotangent of the last hidden state is the sequence output's; the other depths are added to the running
                # cotangent at their layer boundary (run_slice)
                self.d_hidden = d_hidden.to(torch.float32).contiguous()
                top = self.d_hidden[:, -1]
                self.d_seq = top.contiguous() if self.d_seq is None else self.d_seq + top
            self.ws = torch.empty(lib.simx_encoder_workspace_bytes(C.byref(self.cfg), B, S), dtype=torch.uint8,
                                  device=self.dev)
            numel = lib.simx_encoder_grad_numel(C.byref(self.cfg))
            # a sharded gradient exchange (gradsync.FlatGradSync(shard=True)) reduce-scatters equal chunks: room for the tail pad
            self.flat = torch.empty(max(numel, int(getattr(self.module, "_flat_padded_numel", 0))), dtype=torch.float32,
                                    device=self.dev)
        return True

    def run_slice(self, lo: int, hi: int) -> None:
        """backward through layers [lo, hi) (lo == -1: down to and including the embeddings)"""
        if self.flat is None:
            return
        if hi != self.next_hi:
            raise RuntimeError(f"encoder backward slices out of order: expected [.., {self.next_hi}), got [{lo}, {hi})")
        lib = _lib.lib()
        B, S = self.shape
        m = self.module
        if self.embeds:
            if (lo, hi) != (-1, m.enc_cfg.num_layers):
                raise RuntimeError("inputs_embeds passes run their backward in one slice")
            with torch.cuda.device(self.dev):
                stream = torch.cuda.current_stream(self.dev).cuda_stream
                self.d_embeds = torch.empty(B, S, m.enc_cfg.hidden, dtype=torch.float32, device=self.dev)
                _lib.check(lib.simx_encoder_backward_embeds(
                    C.byref(self.cfg), m._param_table(self.params), self.wbuf.data_ptr(), _lib.ptr(self.mask), None, 0, B, S,
                    self.saved.data_ptr(), self.ws.data_ptr(), _lib.ptr(self.d_cls), _lib.ptr(self.d_seq), self.flat.data_ptr(),
                    self.d_embeds.data_ptr(), stream), "simx_encoder_backward_embeds")
            self.next_hi = -1
            self.saved = self.ws = self.d_cls = self.d_seq = None
            return
        with torch.cuda.device(self.dev):
            stream = torch.cuda.current_stream(self.dev).cuda_stream
            table = m._param_table(self.params)

            def layers(a: int, b: int) -> None:
                _lib.check(lib.simx_encoder_backward_layers(
                    C.byref(self.cfg), table, self.wbuf.data_ptr(), self.ids.data_ptr(), _lib.ptr(self.mask),
                    _lib.ptr(self.pos_ids), m.enc_cfg.pad_token_id if self.pos_ids is not None else 0, B, S,
                    self.saved.data_ptr(), self.ws.data_ptr(), _lib.ptr(self.d_cls), _lib.ptr(self.d_seq),
                    self.flat.data_ptr(), a, b, stream), "simx_encoder_backward_layers")

            if self.d_hidden is None:
                layers(lo, hi)
            else:  # one layer at a time: the cotangent of hidden state l joins the running one between layer l and l - 1
                Lh, H = m.enc_cfg.num_layers + 1, m.enc_cfg.hidden
                for l in range(hi - 1, max(lo, 0) - 1, -1):
                    layers(l, l + 1)
                    _lib.check(lib.simx_encoder_add_cotangent(C.byref(self.cfg), self.ws.data_ptr(), B, S,
                                                              self.d_hidden[:, l].data_ptr(), Lh * S * H, stream),
                               "simx_encoder_add_cotangent")
                if lo < 0:
                    layers(-1, 0)
        self.next_hi = lo
        if lo < 0:  # pass complete: release everything but the gradient buffer (kept alive by the views handed out)
            self.saved = self.ws = self.d_cls = self.d_seq = self.d_hidden = None

    def finish(self) -> None:
        """the last slice of this pass has run: hand the flat buffer to the gradient sink, if one is attached"""
        if self.sink is not None and self.flat is not None:
            self.sink(self.module, self)

    def grads_for(self, params, needs) -> tuple:
        if self.flat is None or self.sink is not None:  # with a sink, parameters get their gradients from it, not from autograd
            return (None,) * len(params)
        slices = self.module.grad_slices()
        index = {id(p): i for i, p in enumerate(self.params)}
        out = []
        for p, need in zip(params, needs):
            off, n = slices[index[id(p)]]
            out.append(self.flat[off:off + n].view(p.shape) if need else None)
        return tuple(out)


class _EmbedNode(torch.autograd.Function):
    """bottom of the chain: owns the embedding parameters; its backward finishes the pass"""

    @staticmethod
    def forward(ctx, st: _TowerPass, *params):
        ctx.st, ctx.params = st, params
        return st.token()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _token_grad):
        st: _TowerPass = ctx.st
        if st.flat is not None and st.next_hi > 0:  # layers whose nodes were pruned (nothing trainable between) still run
            st.run_slice(0, st.next_hi)
        st.run_slice(-1, 0)
        st.finish()
        return (None,) + st.grads_for(ctx.params, ctx.needs_input_grad[1:])


class _LayerNode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, st: _TowerPass, layer: int, token, *params):
        ctx.st, ctx.layer, ctx.params = st, layer, params
        return st.token()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _token_grad):
        st: _TowerPass = ctx.st
        l = ctx.layer
        if st.flat is not None:
            st.run_slice(l, st.next_hi)  # normally [l, l+1); wider when the nodes above were pruned
        token_grad = st.token() if ctx.needs_input_grad[2] else None
        if token_grad is None:  # nothing trainable below: this node is the bottom of the chain
            st.finish()
        return (None, None, token_grad) + st.grads_for(ctx.params, ctx.needs_input_grad[3:])


class _TopNode(torch.autograd.Function):
    """hands out the tower's outputs; its backward receives the cotangents and opens the pass"""

    @staticmethod
    def forward(ctx, st: _TowerPass, token, *params):
        ctx.st, ctx.params = st, params
        ctx.set_materialize_grads(False)
        # the pass state must not keep the tensors it hands out: output.grad_fn -> this node -> ctx.st -> output would be a
        # reference cycle, and the saved activations / gradient buffer of every step would wait for the cyclic collector
        cls, seq, hidden = st.cls, st.seq, st.hidden
        st.cls = st.seq = st.hidden = None
        dead = []
        if seq is None:
            seq = st.token()
            dead.append(seq)
        if hidden is None:
            hidden = st.token()
            dead.append(hidden)
        if dead:
            ctx.mark_non_differentiable(*dead)
        return cls, seq, hidden

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_cls, d_seq, d_hidden):
        st: _TowerPass = ctx.st
        live = st.begin_backward(d_cls, d_seq, d_hidden)
        if st.monolithic:  # the whole pass in this node: every parameter is an input here
            if live:
                st.run_slice(-1, st.next_hi)
                st.finish()
            d_in = st.d_embeds if (st.embeds and ctx.needs_input_grad[1]) else None  # `token` slot = inputs_embeds here
            return (None, d_in) + st.grads_for(ctx.params, ctx.needs_input_grad[2:])
        token_grad = st.token() if (live and ctx.needs_input_grad[1]) else None
        # pooler parameters: dead math in the reference (multiplied by 0), dense zero gradients.  Their slice of
        # the flat buffer is zeroed by the first backward slice, which has not run yet — hand out fresh zeros instead.
        grads = tuple(torch.zeros_like(p) if (live and need and st.sink is None) else None
                      for p, need in zip(ctx.params, ctx.needs_input_grad[2:]))
        if live and token_grad is None:  # nothing trainable below the outputs: only the (dead) pooler gets zeros
            st.flat.zero_()
            st.finish()
        return (None, token_grad) + grads
