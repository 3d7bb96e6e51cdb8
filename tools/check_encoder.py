"""GPU bring-up diagnostics for the encoder tower and the loss kernel against the numpy oracle.
Prints per-stage / per-parameter error tables (more useful than a bare assert while bringing kernels up).

    python tools/check_encoder.py [--prec bf16x3|bf16] [--arch bert|distilbert] [--layers 2] [--B 3] [--S 24]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import bert_numpy as bn  # noqa: E402
from oracle import losses_numpy as ln  # noqa: E402
from simxns_b200 import _lib  # noqa: E402
from simxns_b200.encoder import EncoderConfig  # noqa: E402
from simxns_b200.losses import fused_score_loss  # noqa: E402
from simxns_b200.models import HFBertEncoder, HFDistilBertEncoder  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def check_encoder(args):
    prec = _lib.PREC_BF16X3 if args.prec == "bf16x3" else _lib.PREC_BF16
    cfg = bn.bert_config(num_layers=args.layers, hidden=args.hidden, heads=args.hidden // 64, intermediate=args.inter,
                         vocab=args.vocab, max_pos=max(64, args.S), arch=args.arch)
    params = bn.synth_params(cfg, seed=7)
    ids, mask = bn.synth_tokens(args.B, args.S, cfg["vocab"], seed=11, ragged=not args.full)
    ora = bn.EncoderOracle(cfg, params)
    out = ora.forward(ids, mask)
    rng = np.random.RandomState(5)
    d_cls = rng.standard_normal(out["cls"].shape)
    d_seq = rng.standard_normal(out["seq"].shape) * mask[:, :, None]
    ref_grads = ora.backward(d_cls=d_cls, d_seq=d_seq)

    ec = EncoderConfig(arch=args.arch, num_layers=cfg["num_layers"], hidden=cfg["hidden"], heads=cfg["heads"],
                       intermediate=cfg["intermediate"], vocab=cfg["vocab"], max_pos=cfg["max_pos"],
                       type_vocab=cfg["type_vocab"], ln_eps=cfg["ln_eps"])
    cls_ = HFBertEncoder if args.arch == "bert" else HFDistilBertEncoder
    model = cls_(ec, precision=prec)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    model.cuda()
    t_ids, t_mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    t0 = time.time()
    cls, seq = model.encode(t_ids, t_mask, want_sequence=True)
    torch.cuda.synchronize()
    print(f"forward ok in {time.time() - t0:.2f}s")
    bad = 0
    for l in range(cfg["num_layers"] + 1):
        e = rel(cls[:, l].detach().cpu().numpy(), out["cls"][:, l])
        print(f"  cls layer {l}: rel err {e:.3e}")
    valid = mask.astype(bool)
    e = rel(seq.detach().cpu().numpy()[valid], out["seq"][valid])
    print(f"  seq (valid tokens): rel err {e:.3e}")
    tol = 2e-5 if prec == _lib.PREC_BF16X3 else 5e-2
    bad += e > tol
    loss = (cls * torch.from_numpy(d_cls).float().cuda()).sum() + (seq * torch.from_numpy(d_seq).float().cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    worst = 0.0
    # key biases have an exactly-zero true gradient (softmax is shift invariant): floor the scale
    floor = 1e-3 * float(np.median([np.abs(v).max() for v in ref_grads.values() if np.abs(v).max() > 0]))
    for name, p in model.named_parameters():
        g = p.grad.cpu().numpy()
        r = ref_grads[name]
        fl = floor * (30.0 if ("key.bias" in name or "k_lin.bias" in name) else 1.0)
        e = float(np.abs(g - r).max() / max(np.abs(r).max(), fl))
        worst = max(worst, e)
        flag = "" if e <= (1e-4 if prec == _lib.PREC_BF16X3 else 8e-2) else "   <-- BAD"
        if flag or args.verbose:
            print(f"  grad {name:60s} rel err {e:.3e}  |ref|max {np.abs(r).max():.3e}{flag}")
        bad += bool(flag)
    print(f"worst grad rel err {worst:.3e}")
    return bad


def check_loss(args):
    rng = np.random.RandomState(1)
    GQ, group, H = 16, 8, 128
    Q = rng.standard_normal((GQ, H)).astype(np.float32)
    Cm = rng.standard_normal((GQ * group, H)).astype(np.float32)
    t = (3 * rng.standard_normal((GQ, group))).astype(np.float32)
    bad = 0
    for terms, w, T in ((["nll_global"], {}, 1.0), (["kl_local"], {}, 1.0), (["kl_local"], {}, 2.0),
                        (["nll_local", "kd_local"], dict(nll_local=0.1, kd_local=0.9), 4.0),
                        (["nll_local", "ce_local"], dict(nll_local=0.1, ce_local=0.9), 1.0),
                        (["nll_global", "ce_local"], dict(nll_global=0.2, ce_local=1.0), 1.0)):
        ref, rdQ, rdC = ln.score_loss(Q, Cm, terms=terms, group=group, pos_idx=np.arange(GQ) * group, teacher=t,
                                      temperature=T, w=w)
        q = torch.from_numpy(Q).cuda().requires_grad_(True)
        c = torch.from_numpy(Cm).cuda().requires_grad_(True)
        loss, detail = fused_score_loss(q, c, terms, teacher=torch.from_numpy(t).cuda(), weights=w, temperature=T,
                                        group_size=group, gather=False)
        loss.backward()
        el = abs(loss.item() - ref["total"])
        eq, ec = rel(q.grad.cpu().numpy(), rdQ), rel(c.grad.cpu().numpy(), rdC)
        ok = el < 1e-4 and eq < 1e-4 and ec < 1e-4
        print(f"  loss {'+'.join(terms):24s} T={T}: |dloss| {el:.2e}  dQ rel {eq:.2e}  dC rel {ec:.2e}  "
              f"n_correct {detail[6].item():.0f}/{ref['n_correct']:.0f} {'' if ok else '<-- BAD'}")
        bad += not ok
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prec", default="bf16x3")
    ap.add_argument("--arch", default="bert")
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--inter", type=int, default=256)
    ap.add_argument("--vocab", type=int, default=512)
    ap.add_argument("--B", type=int, default=3)
    ap.add_argument("--S", type=int, default=24)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--what", default="encoder,loss")
    args = ap.parse_args()
    bad = 0
    if "loss" in args.what:
        bad += check_loss(args)
    if "encoder" in args.what:
        bad += check_encoder(args)
    print("RESULT", "OK" if bad == 0 else f"{bad} BAD")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
