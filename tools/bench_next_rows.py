#!/usr/bin/env python
"""Measurements of the SURVEY §8(f) "next" rows on one B200 (not the headline metric — that is bench.py):

  encode     N1  corpus encoding for the ANN refresh: BiBertEncoder.body_emb under no_grad, batch 256 x 128
                 (co_training_generate.py:97-121)                                          -> passages/s, model TFLOP/s
  search     N1  exact inner-product top-k over the per-GPU shard of MS-MARCO (8.8 M / 8 = 1.1 M x 768), k = 200 / 1000
                 (co_training_generate.py:359-384, 412-427)                                 -> queries/s, GB/s of the select
  teacher    N2  cross-encoder CE step (Reranker fwd + CrossEntropy + bwd) at [N, 16, 160]  (co_training_marco_train.py:225-245)
  optimizer  N3  clip_grad_norm_ + AdamW over both BERT-base towers (219 M parameters): fused kernels vs torch's own
                 (co_training_marco_train.py:57-69, 246-263)                                -> ms/step, GB/s vs HBM peak

Every number: CUDA events on the launching stream, 3 warm-up + K timed iterations, inputs larger than L2.
One JSON object per workload on stdout; `--out` also writes them to a file (profiles/).
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def timed(fn, steps, warmup=3):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


def fwd_flops(layers, S, H=768, I=3072):
    return layers * (8 * S * H * H + 4 * S * H * I + 4 * S * S * H)


def bench_encode(args):
    import torch

    from simxns_b200 import _lib
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.models import BiBertEncoder, HFBertEncoder
    from simxns_b200.search import embed_passages

    ec = EncoderConfig(num_layers=args.layers)
    torch.manual_seed(0)
    model = BiBertEncoder(None, question_model=HFBertEncoder(ec, precision=_lib.PREC_BF16)).cuda().eval()
    n = 256 * args.encode_batches
    ids = torch.randint(1000, ec.vocab, (n, 128), device="cuda")
    ids[:, 0] = 101
    out = []
    _, tf_peak, src = peaks()
    for bs in (256, 2048):  # the reference's batch, and the batch embed_passages uses by default
        ms = timed(lambda: embed_passages(model.body_emb, ids, bs), args.steps)
        pps = n / (ms * 1e-3)
        tf = pps * fwd_flops(args.layers, 128) / 1e12
        out.append(dict(workload=f"encode (N1): body_emb no_grad, batches of {bs} x 128, bert-base", passages_per_s=pps,
                        ms_per_batch=ms * bs / n, model_tflops=tf, frac_of_sustained_bf16=tf / tf_peak, peak_source=src,
                        reference_note="reference embeds 8.8 M passages per refresh: %.0f s at this rate on one GPU" % (8.8e6 / pps)))
    return out


def bench_search(args):
    import torch

    from simxns_b200 import _lib
    from simxns_b200.search import FlatIPIndex

    N, H, nq = args.search_rows, 768, args.search_queries
    g = torch.Generator(device="cuda").manual_seed(0)
    index = FlatIPIndex(H, capacity=N)
    base = torch.randn(H, device="cuda", generator=g)
    for b in range(0, N, 131072):
        n = min(131072, N - b)
        index.add(0.5 * torch.randn(n, H, device="cuda", generator=g) + base)
    q = 0.5 * torch.randn(nq, H, device="cuda", generator=g) + base
    hbm, tf_peak, src = peaks()
    out = []
    for k in (200, 1000):
        _lib.profile_begin()
        index.search_local(q, k)
        kern = _lib.profile_end()
        ms = timed(lambda: index.search_local(q, k), args.steps)
        sel = next(x for x in kern if x["name"] == "topk_select")
        sc = next(x for x in kern if x["name"] == "topk_scores")
        out.append(dict(workload=f"search (N1): exact IP top-{k}, {nq} queries x {N} x {H} shard (fp32-class scores)",
                        queries_per_s=nq / (ms * 1e-3), ms=ms, pairs_per_s=nq * N / (ms * 1e-3),
                        kernels={x["name"]: dict(launches=x["launches"], ms=x["ms"]) for x in kern},
                        select_gbs=sel["bytes"] / sel["ms"] / 1e6, select_frac_of_hbm=sel["bytes"] / sel["ms"] / 1e6 / hbm,
                        scores_tflops=sc["flops"] / sc["ms"] / 1e9, scores_frac_of_sustained_bf16=sc["flops"] / sc["ms"] / 1e9 / tf_peak,
                        peak_source=src))
    return out


def bench_teacher(args):
    import torch

    from simxns_b200 import _lib
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.models import HFBertEncoder, Reranker

    ec = EncoderConfig(num_layers=args.layers)
    torch.manual_seed(0)
    rr = Reranker(HFBertEncoder(ec, precision=_lib.PREC_BF16), ec.hidden).cuda()
    N, M, L = args.teacher_queries, 16, 160
    ids = torch.randint(1000, ec.vocab, (N, M, L), device="cuda")
    ids[:, :, 0] = 101
    mask = torch.ones_like(ids)
    tgt = torch.zeros(N, dtype=torch.long, device="cuda")

    def step():
        loss = torch.nn.functional.cross_entropy(rr(ids, mask), tgt)  # co_training_marco_train.py:231-233
        loss.backward()
        rr.zero_grad(set_to_none=True)

    ms = timed(step, args.steps)
    pairs = N * M / (ms * 1e-3)
    _, tf_peak, src = peaks()
    tf = pairs * 3 * fwd_flops(args.layers, L) / 1e12
    return dict(workload=f"teacher (N2): Reranker CE step, bert-base {args.layers}L, [{N}, {M}, {L}]", pairs_per_s=pairs,
                ms_per_step=ms, model_tflops=tf, frac_of_sustained_bf16=tf / tf_peak, peak_source=src)


def bench_optimizer(args):
    import torch

    from simxns_b200 import optim as sx
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.models import BiBertEncoder, HFBertEncoder

    ec = EncoderConfig(num_layers=args.layers)

    def make():
        torch.manual_seed(0)
        m = BiBertEncoder(None, question_model=HFBertEncoder(ec), ctx_model=HFBertEncoder(ec)).cuda()
        for p in m.parameters():
            p.grad = torch.randn_like(p) * 1e-3
        no_decay = ["bias", "LayerNorm.weight"]  # get_optimizer, co_training_marco_train.py:57-65
        groups = [dict(params=[p for n, p in m.named_parameters() if not any(nd in n for nd in no_decay)], weight_decay=0.01),
                  dict(params=[p for n, p in m.named_parameters() if any(nd in n for nd in no_decay)], weight_decay=0.0)]
        return m, groups

    m1, g1 = make()
    nparam = sum(p.numel() for p in m1.parameters())
    fused = sx.AdamW(g1, lr=1e-5, eps=1e-8, flavor="hf")
    ms_fused = timed(lambda: fused.step(max_grad_norm=1.0), args.steps)
    ms_fused_2 = timed(lambda: (sx.clip_grad_norm_(m1.parameters(), 1.0), fused.step()), args.steps)
    del fused, m1, g1
    torch.cuda.empty_cache()
    m2, g2 = make()
    stock = torch.optim.AdamW(g2, lr=1e-5, eps=1e-8)

    def stock_step():
        torch.nn.utils.clip_grad_norm_(m2.parameters(), 1.0)
        stock.step()

    ms_stock = timed(stock_step, args.steps)
    hbm, _, src = peaks()
    gbs = nparam * 32.0 / ms_fused / 1e6  # 4 B norm pass + 28 B update per parameter
    return dict(workload=f"optimizer (N3): clip_grad_norm_ + AdamW, {nparam / 1e6:.1f} M fp32 parameters (two bert-base towers)",
                fused_ms=ms_fused, fused_clip_then_step_ms=ms_fused_2, torch_foreach_ms=ms_stock, speedup_vs_torch=ms_stock / ms_fused,
                algorithmic_bytes_per_param=32, achieved_gbs=gbs, frac_of_hbm=gbs / hbm, peak_source=src,
                launches_fused=2, note="torch arm = torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW (foreach) on the same GPU")


def bench_train_full(args):
    """the whole retriever micro-step of the co-training loop INCLUDING the optimizer tail (simxns_b200.train.retriever_step:
    both towers fwd, fused SimANS KL loss against resident teacher scores, bwd, fused clip + AdamW, scheduler, zero_grad)"""
    import types

    import torch

    from simxns_b200 import _lib
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.models import BiBertEncoder, HFBertEncoder
    from simxns_b200.train import get_linear_schedule_with_warmup, get_optimizer, retriever_step

    ec = EncoderConfig(num_layers=args.layers)
    torch.manual_seed(0)
    model = BiBertEncoder(None, question_model=HFBertEncoder(ec, precision=_lib.PREC_BF16),
                          ctx_model=HFBertEncoder(ec, precision=_lib.PREC_BF16)).cuda()
    a = types.SimpleNamespace(temperature_distill=1.0, scale_simmila=False, gradient_accumulation_steps=1, max_grad_norm=1.0,
                              optimizer="adamW")
    opt = get_optimizer(a, model, weight_decay=0.01, lr=1e-5, eps=1e-8)
    sched = get_linear_schedule_with_warmup(opt, 10, 1000)
    nq, group = 128, 16
    q = torch.randint(1000, ec.vocab, (nq, 32), device="cuda")
    p = torch.randint(1000, ec.vocab, (nq * group, 128), device="cuda")
    batch = dict(student=[q, torch.ones_like(q), p, torch.ones_like(p)],
                 teacher=[None, None, None], teacher_scores=3 * torch.randn(nq, group, device="cuda"))
    state = dict(step=0)

    def step():
        state["step"] += 1
        retriever_step(a, model, None, batch, opt, sched, state["step"])

    ms = timed(step, args.steps)
    return dict(workload="train_full: retriever micro-step incl. optimizer (128 q x 16 p, q 32 / p 128, bert-base, SimANS KL loss, "
                         "fused clip + AdamW)", ms_per_step=ms, pairs_per_s=nq * group / (ms * 1e-3),
                note="headline bench.py excludes the optimizer (SURVEY 8d); this is the loop body end to end")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="encode,search,teacher,optimizer,train_full")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--encode-batches", type=int, default=32)
    ap.add_argument("--search-rows", type=int, default=1105228)  # 8 841 823 MS-MARCO passages / 8 GPUs
    ap.add_argument("--search-queries", type=int, default=4096)
    ap.add_argument("--teacher-queries", type=int, default=128)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch

    from simxns_b200 import _lib

    torch.cuda.set_device(0)
    _lib.require_device()
    results = []
    for name in args.only.split(","):
        r = dict(encode=bench_encode, search=bench_search, teacher=bench_teacher, optimizer=bench_optimizer,
                 train_full=bench_train_full)[name](args)
        for x in (r if isinstance(r, list) else [r]):
            print(json.dumps(x), flush=True)
            results.append(x)
        torch.cuda.empty_cache()
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
