"""Where the LEAD ColBERT-teacher step spends its time (one B200): tower passes with / without full hidden states, the
late-interaction scores alone, and the whole step.  CUDA-event timing after warm-up."""
import json
import sys
import types
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from simxns_b200 import _lib, lead_models  # noqa: E402
from simxns_b200.encoder import EncoderConfig  # noqa: E402
from simxns_b200.losses import colbert_score, lead_col_distill_loss  # noqa: E402


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    group, q_len, p_len = 16, 32, 128
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    t_cfg = EncoderConfig(num_layers=12)
    col = lead_models.HFColBertEncoder(t_cfg, precision=_lib.PREC_BF16).to(dev).train()
    g = torch.Generator(device="cpu").manual_seed(1)
    q = torch.randint(1000, t_cfg.vocab, (nq, q_len), generator=g).to(dev)
    p = torch.randint(1000, t_cfg.vocab, (nq * group, p_len), generator=g).to(dev)
    qm, pm = torch.ones_like(q), torch.ones_like(p)
    res = {"queries": nq}

    def towers_cls():
        cls, _ = col.encode(p, pm)
        cls[:, -1].sum().backward()
        col.zero_grad(set_to_none=True)

    def towers_hidden():
        _, seq, hid = col.encode(p, pm, want_sequence=True, want_hidden=True)
        (hid[:, 8].sum() + seq.sum()).backward()
        col.zero_grad(set_to_none=True)

    res["teacher_docs_cls_only_ms"] = timed(towers_cls)
    res["teacher_docs_full_hidden_ms"] = timed(towers_hidden)

    qe = torch.nn.functional.normalize(torch.randn(nq, q_len, 768, device=dev), dim=2).requires_grad_(True)
    de = torch.nn.functional.normalize(torch.randn(nq * group, p_len, 768, device=dev), dim=2).requires_grad_(True)

    def score():
        s = colbert_score(qe, de, pm, nq)
        s.sum().backward()
        qe.grad = de.grad = None

    res["colbert_score_fwd_bwd_ms"] = timed(score)
    res["peak_before_step_gb"] = torch.cuda.max_memory_allocated() / 2**30
    del qe, de
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()

    # the whole step, phase by phase (host clock around synchronised phases: includes launch / Python overhead)
    import time
    s_cfg = EncoderConfig(arch="distilbert", num_layers=6, type_vocab=0)
    db = lead_models.BiBertEncoder(None, question_model=lead_models.HFDistilBertEncoder(s_cfg, precision=_lib.PREC_BF16)).to(dev).train()
    colm = lead_models.BiBertEncoder(types.SimpleNamespace(model_type="colbert"), question_model=col).to(dev).train()
    sel_db, sel_col = [1, 2, 3, 4, 5], [7, 8, 9, 10, 11]
    phases = {"towers_fwd": 0.0, "loss_fwd": 0.0, "backward": 0.0}
    n = 8
    for it in range(n + 3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        q_col, c_col, qh_col, ch_col, _ = colm(q, qm, p, pm)
        q_db, c_db, ql_db, cl_db, _ = db(q, qm, p, pm)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if it == 3:
            res["mem_after_towers_fwd_gb"] = torch.cuda.memory_allocated() / 2**30
        loss, _ = lead_col_distill_loss(q_db, c_db, q_col, c_col, pm, ql_db[:, :, 0], cl_db[:, :, 0], qh_col, ch_col,
                                        selected_db=sel_db, selected_col=sel_col, temperature=1.0, layer_temperature=10.0,
                                        reweight=True, train_col=True, gather=False)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if it == 3:
            res["mem_after_loss_fwd_gb"] = torch.cuda.memory_allocated() / 2**30
            torch.cuda.reset_peak_memory_stats()
        loss.backward()
        torch.cuda.synchronize(); t3 = time.perf_counter()
        if it == 3:
            res["peak_during_backward_gb"] = torch.cuda.max_memory_allocated() / 2**30
        colm.zero_grad(set_to_none=True); db.zero_grad(set_to_none=True)
        if it >= 3:
            phases["towers_fwd"] += (t1 - t0) * 1e3 / n
            phases["loss_fwd"] += (t2 - t1) * 1e3 / n
            phases["backward"] += (t3 - t2) * 1e3 / n
    res["step_phases_ms"] = phases
    res["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2**30
    print(json.dumps(res))


if __name__ == "__main__":
    main()
