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

    def c_struct(self, precision: int) -> _lib.EncoderCfg:
        return _lib.EncoderCfg(
            arch=_lib.ARCH_BERT if self.arch == "bert" else _lib.ARCH_DISTILBERT,
            num_layers=self.num_layers, hidden=self.hidden, heads=self.heads, intermediate=self.intermediate,
            vocab=self.vocab, max_pos=self.max_pos, type_vocab=self.type_vocab if self.arch == "bert" else 0,
            precision=precision, ln_eps=self.ln_eps)

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

    def _c_cfg(self) -> _lib.EncoderCfg:
        return self.enc_cfg.c_struct(self.precision)

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
            cfg = self._c_cfg()
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
            cfg = self._c_cfg()
            n = lib.simx_encoder_num_params(C.byref(cfg))
            assert n == len(self._ordered), (n, len(self._ordered))
            self._grad_slices = [(lib.simx_encoder_grad_offset(C.byref(cfg), i),
                                  lib.simx_encoder_param_numel(C.byref(cfg), i)) for i in range(n)]
        return self._grad_slices

    # ---- compute ------------------------------------------------------------------------------------------
    def encode(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
               want_sequence: bool = False):
        """-> (cls_layers [B, L+1, H] fp32, sequence_output [B,S,H] fp32 or None).
        cls_layers[:, -1] is the reference's pooled_output; cls_layers[:, l+1] the [CLS] row of layer l."""
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, S]")
        params = self._ordered
        cls, seq = _EncoderFn.apply(self, input_ids, attention_mask, bool(want_sequence), *params)
        return cls, (seq if want_sequence else None)


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: B200EncoderBase, ids, mask, want_seq, *params):
        lib = _lib.lib()
        module._check_params(params)
        dev = params[0].device
        if ids.device != dev:
            raise RuntimeError(f"input_ids on {ids.device}, parameters on {dev}")
        with torch.cuda.device(dev):
            _lib.require_device()
            ids = ids.to(torch.int64).contiguous()
            if mask is not None:
                mask = mask.to(device=dev, dtype=torch.int64).contiguous()
                if mask.shape != ids.shape:
                    raise ValueError("attention_mask must have the shape of input_ids")
            B, S = ids.shape
            cfg = module._c_cfg()
            ec = module.enc_cfg
            if S > min(ec.max_pos, 512):
                raise ValueError(f"sequence length {S} exceeds min(max_position_embeddings={ec.max_pos}, 512)")
            pos_ids = None
            if ec.roberta_positions:
                # transformers create_position_ids_from_input_ids: pad + running count of real tokens, pad for padding
                if S + ec.pad_token_id >= ec.max_pos:
                    raise ValueError(f"sequence length {S} needs position {S + ec.pad_token_id} >= "
                                     f"max_position_embeddings={ec.max_pos}")
                real = (ids != ec.pad_token_id).to(torch.int64)
                pos_ids = (torch.cumsum(real, dim=1) * real + ec.pad_token_id).contiguous()
            needs_grad = any(ctx.needs_input_grad[4:])  # grad mode is always off inside Function.forward
            wbuf = module._weights(params)
            table = module._param_table(params)
            stream = torch.cuda.current_stream(dev).cuda_stream
            cls = torch.empty(B, ec.num_layers + 1, ec.hidden, dtype=torch.float32, device=dev)
            seq = torch.empty(B, S, ec.hidden, dtype=torch.float32, device=dev) if want_seq else None
            ws = torch.empty(lib.simx_encoder_workspace_bytes(C.byref(cfg), B, S), dtype=torch.uint8, device=dev)
            saved = None
            if needs_grad:
                saved = torch.empty(lib.simx_encoder_saved_bytes(C.byref(cfg), B, S), dtype=torch.uint8, device=dev)
            if pos_ids is None:
                _lib.check(lib.simx_encoder_forward(C.byref(cfg), table, wbuf.data_ptr(), ids.data_ptr(), _lib.ptr(mask), B, S,
                                                    _lib.ptr(saved), ws.data_ptr(), cls.data_ptr(), _lib.ptr(seq), stream),
                           "simx_encoder_forward")
            else:
                _lib.check(lib.simx_encoder_forward_pos(C.byref(cfg), table, wbuf.data_ptr(), ids.data_ptr(), _lib.ptr(mask),
                                                        pos_ids.data_ptr(), B, S, _lib.ptr(saved), ws.data_ptr(),
                                                        cls.data_ptr(), _lib.ptr(seq), stream), "simx_encoder_forward_pos")
        ctx.module = module
        ctx.saved_act = saved
        ctx.ids, ctx.mask, ctx.shape, ctx.pos_ids = ids, mask, (B, S), pos_ids
        ctx.want_seq = want_seq
        ctx.wbuf = wbuf
        ctx.params = params
        ctx.set_materialize_grads(False)
        if seq is None:
            seq = cls.new_empty(0)
            ctx.mark_non_differentiable(seq)
        return cls, seq

    @staticmethod
    def backward(ctx, d_cls, d_seq):
        module: B200EncoderBase = ctx.module
        params = ctx.params
        n_lead = 4
        if ctx.saved_act is None:
            raise RuntimeError("encoder backward without saved activations (forward ran under no_grad)")
        if d_cls is None and (d_seq is None or not ctx.want_seq):
            return (None,) * (n_lead + len(params))
        lib = _lib.lib()
        dev = params[0].device
        B, S = ctx.shape
        with torch.cuda.device(dev):
            cfg = module._c_cfg()
            stream = torch.cuda.current_stream(dev).cuda_stream
            if d_cls is not None:
                d_cls = d_cls.to(torch.float32).contiguous()
            if d_seq is not None and ctx.want_seq:
                d_seq = d_seq.to(torch.float32).contiguous()
            else:
                d_seq = None
            ws = torch.empty(lib.simx_encoder_workspace_bytes(C.byref(cfg), B, S), dtype=torch.uint8, device=dev)
            flat = torch.empty(lib.simx_encoder_grad_numel(C.byref(cfg)), dtype=torch.float32, device=dev)
            if ctx.pos_ids is None:
                _lib.check(lib.simx_encoder_backward(C.byref(cfg), module._param_table(params), ctx.wbuf.data_ptr(),
                                                     ctx.ids.data_ptr(), _lib.ptr(ctx.mask), B, S, ctx.saved_act.data_ptr(),
                                                     ws.data_ptr(), _lib.ptr(d_cls), _lib.ptr(d_seq), flat.data_ptr(), stream),
                           "simx_encoder_backward")
            else:
                _lib.check(lib.simx_encoder_backward_pos(C.byref(cfg), module._param_table(params), ctx.wbuf.data_ptr(),
                                                         ctx.ids.data_ptr(), _lib.ptr(ctx.mask), ctx.pos_ids.data_ptr(),
                                                         module.enc_cfg.pad_token_id, B, S, ctx.saved_act.data_ptr(),
                                                         ws.data_ptr(), _lib.ptr(d_cls), _lib.ptr(d_seq), flat.data_ptr(),
                                                         stream), "simx_encoder_backward_pos")
        ctx.saved_act = None
        grads = []
        for p, (off, n), need in zip(params, module.grad_slices(), ctx.needs_input_grad[n_lead:]):
            grads.append(flat[off:off + n].view(p.shape) if need else None)
        return (None,) * n_lead + tuple(grads)
