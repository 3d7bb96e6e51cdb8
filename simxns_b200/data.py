"""Batch layout of the SimANS MS-MARCO loop: what `Rocketqa_v2Dataset.__getitem__` does AFTER tokenising
(/root/reference/SimANS/utils/MARCO_until_new.py:216-233) and its collate function (:240-260).  Pure host logic on token
ids (the HF tokenizer itself is out of scope); every tensor is int64 on the CPU exactly as the reference's DataLoader
workers produce them, so the dict drops into `train()` / `simxns_b200.train.retriever_step` unchanged.

    q_ids, ctx_ids, ce_ids = pack_example(question_token_ids, ctx_token_ids, pad_token_id, sep_token_id)
    batch = collate([ (q_ids, ctx_ids, ce_ids), ... ])
    batch == {'student': [q [B,32], q_mask, d [B*(1+n),128], d_mask, positive_ctx_indices],
              'teacher': [ce [B,1+n,160], ce_mask, tgt [B,1+n]]}

`collate_wiki` is the NQ / TriviaQA variant (/root/reference/SimANS/utils/util_wiki.py:665-698): ragged token lists padded
to the BATCH maximum (queries are not truncated there), keys 'retriever' / 'reranker' / 'answers' /
'reranker_ctx_start_end' — the reason the kernels take any S <= 512.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def pack_example(question_token_ids: Sequence[int], ctx_token_ids: Sequence[Sequence[int]], pad_token_id: int = 0,
                 sep_token_id: int = 102, max_query_length: int = 32, max_seq_length: int = 128,
                 max_ce_length: int = 160) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """MARCO_until_new.py:216-233.  `question_token_ids` / `ctx_token_ids[j]` are the tokenizer's outputs WITH special
    tokens ([CLS] q [SEP]; [CLS] title [SEP] text [SEP]), already truncated to 32 / 128.  The cross-encoder input of a
    pair is the query followed by the passage without its [CLS] (and without its last token when that is [SEP]),
    zero-... pad-padded to 160; positive first."""

    def remove_special_token(token_id):
        if token_id[-1] == sep_token_id:
            return list(token_id[1:-1])
        return list(token_id[1:])

    q = list(question_token_ids)
    ctx = [list(c) for c in ctx_token_ids]
    c_e = [q + remove_special_token(c) for c in ctx]
    if len(q) > max_query_length or any(len(c) > max_seq_length for c in ctx) or any(len(t) > max_ce_length for t in c_e):
        raise ValueError("token lists longer than the padded lengths: truncate in the tokenizer (max_length=32 / 128)")
    q_t = torch.LongTensor(q + [pad_token_id] * (max_query_length - len(q)))
    ctx_t = torch.LongTensor([c + [pad_token_id] * (max_seq_length - len(c)) for c in ctx])
    ce_t = torch.LongTensor([t + [pad_token_id] * (max_ce_length - len(t)) for t in c_e])
    return q_t, ctx_t, ce_t


def collate(features: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], dtype: torch.dtype = torch.int64) -> dict:
    """Rocketqa_v2Dataset.get_collate_fn (MARCO_until_new.py:240-260): passages of query i occupy rows
    i*(1+n) .. i*(1+n)+n of the student's passage tensor, positive first; masks are `ids != 0`.
    dtype=torch.int32 halves the bytes of every id / mask tensor on the host->device wire (the kernels read either width:
    simx_encoder_cfg.id_bytes); the default is the reference's int64."""
    doc_per_question = features[0][1].size(0)
    q_tensor = torch.stack([feature[0] for feature in features], dim=0)
    doc_tensor = torch.cat([feature[1] for feature in features])
    ctx_tensor_out = torch.cat([feature[2] for feature in features])
    positive_ctx_indices = [i * doc_per_question for i in range(len(features))]
    q_num, d_num = q_tensor.size(0), doc_tensor.size(0)
    tgt_tensor = torch.zeros((d_num), dtype=torch.long)
    tgt_tensor[positive_ctx_indices] = 1
    ctx_tensor_out = ctx_tensor_out.reshape(q_num, d_num // q_num, -1)
    tgt_tensor = tgt_tensor.reshape(q_num, d_num // q_num)
    c = lambda t: t.to(dtype)
    return {'student': [c(q_tensor), c(q_tensor != 0), c(doc_tensor), c(doc_tensor != 0), positive_ctx_indices],
            'teacher': [c(ctx_tensor_out), c(ctx_tensor_out != 0), tgt_tensor]}


def collate_wiki(features) -> dict:
    """TraditionDataset.get_collate_fn (util_wiki.py:665-698).  feature = (question_token_ids, ctx_token_ids,
    c_e_token_ids, answers, start_end) with plain Python lists of token ids."""
    q_list, d_list, c_e_input_list, positive_ctx_indices, c_e_ctx_start_end = [], [], [], [], []
    for feature in features:
        positive_ctx_indices.append(len(d_list))
        q_list.append(feature[0])
        d_list.extend(feature[1])
        c_e_input_list.extend(feature[2])
        c_e_ctx_start_end.append(feature[4])
    max_q_len = max(len(q) for q in q_list)
    max_d_len = max(len(d) for d in d_list)
    max_c_e_len = max(len(d) for d in c_e_input_list)
    q_tensor = torch.LongTensor([q + [0] * (max_q_len - len(q)) for q in q_list])
    doc_tensor = torch.LongTensor([d + [0] * (max_d_len - len(d)) for d in d_list])
    ctx_tensor_out = torch.LongTensor([c + [0] * (max_c_e_len - len(c)) for c in c_e_input_list])
    q_num, d_num = len(q_list), len(d_list)
    tgt_tensor = torch.zeros((len(c_e_input_list)), dtype=torch.long)
    tgt_tensor[positive_ctx_indices] = 1
    ctx_tensor_out = ctx_tensor_out.reshape(q_num, d_num // q_num, -1)
    tgt_tensor = tgt_tensor.reshape(q_num, d_num // q_num)
    return {'reranker': [ctx_tensor_out, (ctx_tensor_out != 0).long(), tgt_tensor],
            'retriever': [q_tensor, (q_tensor != 0).long(), doc_tensor, (doc_tensor != 0).long(), positive_ctx_indices],
            "answers": [feature[3] for feature in features],
            "reranker_ctx_start_end": c_e_ctx_start_end}
