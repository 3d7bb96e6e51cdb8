"""The fair GPU bar: the reference's own model code path — HF BertModel wrapper (oracle/ref_torch.py port) +
PyTorch eager + cuBLAS/SDPA — running the same train step on the same B200, fp32 and bf16-autocast.
Not part of bench.py's contract; run on the GPU box:   python tools/ref_gpu_bar.py [--queries 32]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import bert_base_cfg  # noqa: E402
from oracle.ref_torch import BiBertEncoderRef, nll_global_loss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=32)
    ap.add_argument("--group", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    cfg = bert_base_cfg()
    torch.manual_seed(0)
    model = BiBertEncoderRef(cfg).cuda().train()
    g = torch.Generator().manual_seed(1234)
    q = torch.randint(1000, cfg["vocab"], (a.queries, 32), generator=g).cuda()
    p = torch.randint(1000, cfg["vocab"], (a.queries * a.group, 128), generator=g).cuda()
    qm, pm = torch.ones_like(q), torch.ones_like(p)
    out = {}
    for mode in ("fp32", "tf32", "bf16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = mode == "tf32"

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
                qe, ce = model(q, qm, p, pm)
            loss = nll_global_loss(qe.float(), ce.float(), a.group) if False else None
            scores = qe.float() @ ce.float().t()
            tgt = torch.arange(0, a.queries * a.group, a.group, device="cuda")
            loss = torch.nn.functional.cross_entropy(scores, tgt)
            loss.backward()
            model.zero_grad(set_to_none=True)

        try:
            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            out[mode] = dict(ms_per_step=ms, pairs_per_s=a.queries * a.group / ms * 1e3,
                             peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30)
        except torch.cuda.OutOfMemoryError as e:
            out[mode] = dict(error="OOM")
            torch.cuda.empty_cache()
    print(json.dumps(dict(what="reference model code (HF BertModel eager) on 1x B200", queries=a.queries, group=a.group, **out)))


if __name__ == "__main__":
    main()
