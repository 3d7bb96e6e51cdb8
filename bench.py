#!/usr/bin/env python
"""Headline benchmark: dual-encoder student train step (fwd + contrastive loss + bwd, no optimizer, no data loading —
SURVEY §8d) in query+passage pairs/s at the MS-MARCO shape of BASELINE.json (BERT-base, q_len 32 / p_len 128,
1 positive + 15 hard negatives per query, 128 queries per GPU = the per-GPU shard of configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo (libsimx_b200 on B200)
    torchrun ... bench.py --gpus N --steps K --warmup W      # N > 1: one rank per GPU, NCCL
    python bench.py --impl reference ...                     # the reference's own CPU path, all host threads

One step = BiBertEncoder forward of both towers, all-gather of the [CLS] embeddings over the N ranks, the in-batch
InfoNCE over the global [N*128, N*2048] score matrix, backward through both towers (and, for N > 1, DDP's gradient
all-reduce).  Weak scaling: per-GPU work is fixed.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "query+passage pairs/sec (seq 32/128, 1+15 neg)"
UNIT = "pairs/s"
# algorithmic FLOPs of one (query, passage) pair, fwd+bwd, BERT-base 1 q (32) + 16 p (128) group / 16
GFLOP_PER_PAIR = 68.07


def bert_base_cfg(layers=12):
    return dict(arch="bert", num_layers=layers, hidden=768, heads=12, intermediate=3072, vocab=30522, max_pos=512,
                type_vocab=2, ln_eps=1e-12)


def flops_per_pair(layers, q_len, p_len, group, H=768, I=3072):
    seq = lambda S: layers * (8 * S * H * H + 4 * S * H * I + 4 * S * S * H)
    return 3.0 * (seq(q_len) + group * seq(p_len)) / group


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(tflops_burst=d.get("bf16_tflops"), tflops_sustained=d.get("bf16_tflops_sustained"),
                    hbm_gbs=d.get("hbm_gbs"), source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        # median over the samples taken under load (upper half of the sorted clocks excludes idle edges)
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None),
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """reference arm: the reference's CPU implementation of the same step (oracle/ref_torch.py port over the installed
    transformers BertModel), all host threads, bounded sample of the workload"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.ref_torch import time_cpu_step

    nq = args.ref_queries
    r = time_cpu_step(bert_base_cfg(args.layers), nq, args.group, args.q_len, args.p_len, steps=args.steps,
                      warmup=args.warmup)
    sample = (f"{nq} queries x {args.group} passages per step (q_len {args.q_len}, p_len {args.p_len}), fp32, "
              f"{args.steps} steps after {args.warmup} warm-up, in-batch InfoNCE")
    line = dict(impl="reference", metric=METRIC, value=r["pairs_per_s"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=r["s_per_step"] * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="fp32", data="synthetic",
                config=workload_config(args, nq),
                cpu_baseline=dict(value=r["pairs_per_s"], unit=UNIT, cores=r["cores"], kind="port", sample=sample),
                e2e=dict(value=r["pairs_per_s"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, nq):
    return dict(workload=f"SimANS MS-MARCO-Passage shape: bert-base ({args.layers}L), {nq} queries/GPU x (1 pos + "
                         f"{args.group - 1} hard-neg), q_len {args.q_len}, p_len {args.p_len}, in-batch InfoNCE over "
                         f"all-gathered [CLS] embeddings (per-GPU shard of BASELINE configs[2])",
                queries_per_gpu=nq, passages_per_query=args.group, q_len=args.q_len, p_len=args.p_len,
                parallelism=f"dp{args.gpus}", l2="inputs_larger_than_L2 (activations ~78 GB/step)")


def run_parity_mode(args, ec, dev, batch):
    """the same step in SIMX_PREC_BF16X3 — the mode the north_star tolerances (1e-3 / 1e-4) are asserted in.  Three planes
    per activation do not fit 180 GB at this batch, so the towers keep layer inputs only (gradient checkpointing)."""
    import torch

    from simxns_b200 import _lib
    from simxns_b200.losses import fused_score_loss
    from simxns_b200.models import BiBertEncoder, HFBertEncoder

    try:
        torch.manual_seed(0)
        model = BiBertEncoder(None, question_model=HFBertEncoder(ec, precision=_lib.PREC_BF16X3),
                              ctx_model=HFBertEncoder(ec, precision=_lib.PREC_BF16X3)).to(dev)
        for tower in (model.question_model, model.ctx_model):
            tower.gradient_checkpointing_enable()
        b = tuple(t.to(dev) for t in batch)

        def step():
            q, c = model(*b)
            loss, _ = fused_score_loss(q, c, ["nll_global"], group_size=args.group, gather=False)
            loss.backward()
            model.zero_grad(set_to_none=True)

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 2
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        return dict(precision="bf16x3 (fp32-class: 3 bf16 planes, 6 MMA passes)", value=args.queries * args.group / ms * 1e3,
                    unit=UNIT, ms_per_step=ms, steps=n, save_policy="layer inputs only (+1 forward of FLOPs)",
                    peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    except Exception as e:  # the headline must not die with the extras
        return dict(error=f"{type(e).__name__}: {e}"[:300])


def run_gpu_bar(args, dev, batch):
    """the fair bar: the reference's model code path (HF BertModel wrapper, PyTorch eager + cuBLAS / SDPA) running the same
    step on this GPU, fp32 (its default) and under bf16 autocast — builder-side context for the CPU reference arm"""
    import torch

    out = {}
    try:
        from oracle.ref_torch import BiBertEncoderRef

        torch.manual_seed(0)
        model = BiBertEncoderRef(bert_base_cfg(args.layers)).to(dev).train()
        b = tuple(t.to(dev) for t in batch)
        tgt = torch.arange(0, args.queries * args.group, args.group, device=dev)
        for mode in ("fp32", "bf16_autocast"):
            try:
                def step():
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
                        qe, ce = model(*b)
                    loss = torch.nn.functional.cross_entropy(qe.float() @ ce.float().t(), tgt)
                    loss.backward()
                    model.zero_grad(set_to_none=True)

                torch.cuda.reset_peak_memory_stats()
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = 3
                for _ in range(n):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                out[mode] = dict(value=args.queries * args.group / ms * 1e3, unit=UNIT, ms_per_step=ms,
                                 peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
            except torch.cuda.OutOfMemoryError:
                out[mode] = dict(error="OOM")
            model.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
        out["what"] = "reference model code (transformers BertModel, PyTorch eager) on the same GPU, dropout 0"
    except Exception as e:
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from simxns_b200 import _lib
    from simxns_b200.encoder import EncoderConfig
    from simxns_b200.losses import fused_score_loss
    from simxns_b200.models import BiBertEncoder, HFBertEncoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE is {world})")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")  # both levels print NCCL's version banner on stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()

    ec = EncoderConfig(num_layers=args.layers)
    torch.manual_seed(0)
    model = BiBertEncoder(None, question_model=HFBertEncoder(ec, precision=_lib.PREC_BF16),
                          ctx_model=HFBertEncoder(ec, precision=_lib.PREC_BF16)).to(dev)
    step_model = model
    overlap_desc = "query tower on a second CUDA stream" if model.overlap_towers else "off"
    exchange = "none"
    if world > 1 and os.environ.get("SIMX_BENCH_DDP", "0") == "1":
        # the reference's wrap (co_training_marco_train.py:109-114) plus bucket views: .grad aliases the all-reduce
        # buckets, so the 876 MB of gradients are not copied back out of them (2 GPUs: 165.3 -> 159.1 ms per step)
        step_model = torch.nn.parallel.DistributedDataParallel(
            model, device_ids=[local], find_unused_parameters=False, gradient_as_bucket_view=True,
            bucket_cap_mb=int(os.environ.get("SIMX_DDP_BUCKET_MB", "25")))
        exchange = "torch DistributedDataParallel (25 MB buckets)"
    elif world > 1:
        # same contract as the DDP wrap, laid out for these towers: one in-place fp32 all-reduce (AVG) of each tower's flat
        # gradient buffer, issued when that tower's backward completes (simxns_b200/gradsync.py)
        from simxns_b200.gradsync import FlatGradSync

        step_model = FlatGradSync(model, compress=os.environ.get("SIMX_GRAD_COMPRESS") or None)
        exchange = "FlatGradSync: one fp32 NCCL all-reduce per tower over its flat gradient buffer"

    nq, group = args.queries, args.group
    gen = torch.Generator().manual_seed(1234 + rank)
    n_batches = 4
    host = []
    for _ in range(n_batches):
        q = torch.randint(1000, ec.vocab, (nq, args.q_len), generator=gen)
        p = torch.randint(1000, ec.vocab, (nq * group, args.p_len), generator=gen)
        q[:, 0] = 101
        p[:, 0] = 101
        host.append(tuple(t.pin_memory() for t in (q, torch.ones_like(q), p, torch.ones_like(p))))
    resident = [tuple(t.to(dev) for t in b) for b in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def step(batch):
        q, c = step_model(*batch)
        loss, _ = fused_score_loss(q, c, ["nll_global"], group_size=group, gather=True)
        loss.backward()
        model.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------- device-resident timing (the `value`) ----------------
    for i in range(args.warmup):
        step(resident[i % n_batches])
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(resident[i % n_batches])
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    pairs_step = nq * group * world
    value = pairs_step / (ms_step * 1e-3)

    # ---------------- end to end: host buffers in, loss out, every step ----------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = 0.0
    for i in range(args.steps):
        batch = tuple(t.to(dev, non_blocking=True) for t in host[i % n_batches])
        last = step(batch).item()  # device -> host read of the step's loss
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e = dict(value=pairs_step / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d_bytes * world,
               d2h_bytes_per_step=4 * world, ms_per_step=e2e_ms, last_loss=last)

    # ---------------- N > 1: what each GPU needs for the same step WITHOUT any cross-rank exchange ----------------
    # A synchronous data-parallel step runs at the pace of the slowest GPU; under the 1 kW cap the eight GPUs of a box differ
    # by several percent.  step(N) - max(per-rank compute) is the exposed cost of the exchange itself.
    rank_compute = None
    if world > 1 and hasattr(step_model, "detach_sinks"):
        step_model.detach_sinks()

        def local_step(batch):
            q, c = model(*batch)
            loss, _ = fused_score_loss(q, c, ["nll_global"], group_size=group, gather=False)
            loss.backward()
            model.zero_grad(set_to_none=True)

        for i in range(2):
            local_step(resident[i % n_batches])
        barrier()
        n_local = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_local):
            local_step(resident[i % n_batches])
        e1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([e0.elapsed_time(e1) / n_local], device=dev, dtype=torch.float64)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = [round(float(t.item()), 2) for t in every]
        rank_compute = dict(ms_per_step_by_rank=per_rank, slowest=max(per_rank), fastest=min(per_rank),
                            exchange_exposed_ms=round(ms_step - max(per_rank), 2),
                            note="same step per rank with no all-gather / gradient exchange (5 steps); the synchronous step "
                                 "cannot beat `slowest`")
        step_model.attach_sinks()
        barrier()

    # ---------------- the reference's training configuration: dropout 0.1 at every HF site ----------------
    # (SimANS/model/models.py:66-72.  The headline above runs dropout-free like the reference arm's CPU port and the parity
    # fixtures; this is the same step with the Philox masks generated inside the attention kernels / GEMM epilogues.)
    dropout = None
    if not args.no_dropout_run:
        for tower in (model.question_model, model.ctx_model):
            tower.hidden_dropout_prob = tower.attention_probs_dropout_prob = 0.1
        model.train()
        for i in range(2):
            step(resident[i % n_batches])
        barrier()
        d_steps = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(d_steps):
            step(resident[i % n_batches])
        e1.record()
        barrier()
        d_ms = max_over_ranks(e0.elapsed_time(e1)) / d_steps
        dropout = dict(p_hidden=0.1, p_attn=0.1, value=pairs_step / (d_ms * 1e-3), unit=UNIT, ms_per_step=d_ms, steps=d_steps,
                       note="same step with nn.Dropout(0.1) semantics at the embedding / attention-probability / "
                            "dense-output sites, masks regenerated in backward (csrc/dropout.cuh)")
        for tower in (model.question_model, model.ctx_model):
            tower.hidden_dropout_prob = tower.attention_probs_dropout_prob = 0.0

    # ---------------- per-kernel device times of one more step (CUDA events around every launch) ----------------
    kernels, roofline, launches_per_step = None, None, None
    if not args.no_profile:
        barrier()
        prof_steps = 6  # several steps: one step's per-kernel times swing +-10 % with the power-capped clock
        overlap = model.overlap_towers
        model.overlap_towers = False  # per-kernel times of kernels that run alone (the timed steps overlap the two towers)
        torch.cuda.synchronize()
        _lib.profile_begin()
        for i in range(prof_steps):
            step(resident[i % n_batches])
        kernels = _lib.profile_end()
        model.overlap_towers = overlap
        for k in kernels:
            k["launches"] //= prof_steps
            k["ms"] /= prof_steps
            k["flops"] /= prof_steps
            k["bytes"] /= prof_steps
        launches_per_step = sum(k["launches"] for k in kernels)
        peaks = measured_peaks()
        tot = sum(k["ms"] for k in kernels)
        for k in kernels:
            k["share"] = k["ms"] / tot if tot else 0.0
            k["tflops"] = k["flops"] / k["ms"] / 1e9 if k["ms"] > 0 else 0.0
            k["gbs"] = k["bytes"] / k["ms"] / 1e6 if k["ms"] > 0 else 0.0
        gemms = [k for k in kernels if k["flops"] > 0 and k["name"] not in ("attn_fwd", "attn_bwd")]
        dom = max(gemms, key=lambda k: k["ms"]) if gemms else None
        if dom:
            peak = peaks["tflops_sustained"]
            # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (same shape only)
            traffic, tsrc = None, None
            tf = ROOT / "profiles" / "ncu_traffic.json"
            if tf.exists() and (nq, group, args.p_len) == (128, 16, 128):
                t = json.loads(tf.read_text()).get(dom["name"])
                if t:
                    traffic, tsrc = t["dram_bytes_per_launch"], t["source"] + " (passage-tower launch, " + t["shape"] + ")"
            roofline = dict(kernel=dom["name"] + " (gemm_kernel, tcgen05)", bound="tensor", achieved=dom["tflops"],
                            peak=peak, unit="TFLOP/s", frac=dom["tflops"] / peak, traffic=traffic, traffic_source=tsrc,
                            peak_source=peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                            launches=dom["launches"], avg_launch_ms=dom["ms"] / dom["launches"],
                            step_share=dom["share"],
                            all_gemm_tflops=sum(k["flops"] for k in gemms) / sum(k["ms"] for k in gemms) / 1e9,
                            all_gemm_step_share=sum(k["ms"] for k in gemms) / tot)
            wp = ROOT / "profiles" / "r02_hbm_write_probe.json"
            if dom["name"] == "ffn1_fwd" and wp.exists():
                # the FFN-up GEMM stores gelu and gelu' ([tokens, I] bf16 each): its second bound is the WRITE side of HBM
                # (pure stores peak well below the copy figure; DESIGN §5 "What bounds ffn1_fwd")
                tokens = nq * args.q_len + nq * group * args.p_len
                written = 2.0 * tokens * 3072 * 2 * args.layers  # bytes per step
                wpeak = json.loads(wp.read_text())["write_gbs_fill"]
                roofline["store_path"] = dict(written_bytes_per_step=written, achieved_gbs=written / dom["ms"] / 1e6,
                                              peak_gbs=wpeak, frac=written / dom["ms"] / 1e6 / wpeak,
                                              peak_source="profiles/r02_hbm_write_probe.json (torch fill_ over 4 GiB)")

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.ref_torch import time_cpu_step

        r = time_cpu_step(bert_base_cfg(args.layers), args.ref_queries, group, args.q_len, args.p_len, steps=2, warmup=1)
        cpu_baseline = dict(value=r["pairs_per_s"], unit=UNIT, cores=r["cores"], kind="port",
                            sample=f"{args.ref_queries} queries x {group} passages per step, fp32, 2 steps after 1 "
                                   f"warm-up ({r['s_per_step']:.2f} s/step), reference wrapper over transformers "
                                   f"BertModel on the host CPU")

    # ---------------- N = 1 only: the parity mode's speed, and the reference's own GPU code path on this GPU ----------------
    parity_mode, gpu_bar = None, None
    if world == 1 and not args.no_extras:
        del step_model, model, resident
        torch.cuda.empty_cache()
        parity_mode = run_parity_mode(args, ec, dev, host[0])
        torch.cuda.empty_cache()
        gpu_bar = run_gpu_bar(args, dev, host[0])
        torch.cuda.empty_cache()

    if rank == 0:
        fpp = flops_per_pair(args.layers, args.q_len, args.p_len, group)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=dict(workload_config(args, nq), gradient_exchange=exchange,
                                                    tower_overlap=overlap_desc), e2e=e2e,
                    gpu_launches=(launches_per_step * args.steps if launches_per_step else None), clocks=clocks,
                    roofline=roofline, cpu_baseline=cpu_baseline, dropout=dropout, rank_compute=rank_compute,
                    parity_mode=parity_mode,
                    gpu_bar=gpu_bar, model_tflops_per_gpu=value / world * fpp / 1e12, kernels=kernels)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=128, help="queries per GPU")
    ap.add_argument("--group", type=int, default=16, help="passages per query (1 positive + n hard negatives)")
    ap.add_argument("--q-len", type=int, default=32)
    ap.add_argument("--p-len", type=int, default=128)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--ref-queries", type=int, default=2, help="queries per step of the bounded CPU sample")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-dropout-run", action="store_true", help="skip the extra timed run with dropout 0.1")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity-mode and reference-on-GPU runs (N = 1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
