"""GPU bring-up check for the tcgen05 GEMM: every case runs in its own subprocess under a timeout (a hung
mbarrier protocol must not take the box down), compares against torch fp32 matmul, appends to a JSON log.

    python tools/check_gemm.py            # run all cases
    python tools/check_gemm.py --case i   # (internal) run case i in this process
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

# (name, M, N, K, mn_major, planes, epilogue, ksplit)
CASES = [
    ("kk_tiny", 128, 256, 64, 0, 1, 0, 1),
    ("kk_small", 256, 256, 128, 0, 1, 0, 1),
    ("kk_ragged_m", 1000, 768, 768, 0, 1, 0, 1),
    ("kk_n64", 300, 64, 192, 0, 1, 0, 1),
    ("kk_n128", 300, 128, 320, 0, 1, 0, 1),
    ("kk_n_tail", 520, 328, 200, 0, 1, 0, 1),
    ("kk_bias", 512, 2304, 768, 0, 1, 1, 1),
    ("kk_bias_gelu", 512, 3072, 768, 0, 1, 2, 1),
    ("kk_bias_add", 512, 768, 3072, 0, 1, 3, 1),
    ("kk_mul", 512, 3072, 768, 0, 1, 4, 1),
    ("kk_add", 512, 768, 2304, 0, 1, 5, 1),
    ("kk_atomic", 512, 768, 768, 0, 1, 6, 3),
    ("kk_big", 8192, 3072, 768, 0, 1, 2, 1),
    ("mn_tiny", 128, 256, 64, 1, 1, 6, 1),
    ("mn_wgrad", 2304, 768, 4096, 1, 1, 6, 1),
    ("mn_wgrad_split", 768, 3072, 8192, 1, 1, 6, 8),
    ("mn_ragged", 768, 768, 1000, 1, 1, 6, 4),
    ("mn_n64", 192, 64, 520, 1, 1, 6, 2),
    ("x3_kk", 384, 768, 768, 0, 3, 1, 1),
    ("x3_kk_gelu", 256, 3072, 768, 0, 3, 2, 1),
    ("x3_mn", 768, 768, 2048, 1, 3, 6, 4),
]


def split3(x):
    import torch

    hi = x.to(torch.bfloat16)
    r = x - hi.float()
    mid = r.to(torch.bfloat16)
    lo = (r - mid.float()).to(torch.bfloat16)
    return torch.stack([hi, mid, lo])


def run_case(i):
    import torch

    name, M, N, K, mn, planes, epi, ksplit = CASES[i]
    lib = C.CDLL(str(ROOT / "simxns_b200" / "libsimx_b200.so"))
    lib.simx_last_error.restype = C.c_char_p
    P, I, L = C.c_void_p, C.c_int, C.c_int64
    lib.simx_gemm.argtypes = [I, I, I, P, L, L, I, P, L, L, I, I, I, P, P, L, L, P, P, L, L, P, L, I, P]
    torch.manual_seed(i)
    dev = "cuda"
    A32 = torch.randn(M, K, device=dev) * (0.05 if epi == 2 else 1.0)
    B32 = torch.randn(N, K, device=dev)
    bias = torch.randn(N, device=dev)
    aux32 = torch.randn(M, N, device=dev)
    if planes == 1:
        A = A32.to(torch.bfloat16)[None]
        B = B32.to(torch.bfloat16)[None]
        aux = aux32.to(torch.bfloat16)[None]
    else:
        A, B, aux = split3(A32), split3(B32), split3(aux32)
    Af, Bf, auxf = A.float().sum(0), B.float().sum(0), aux.float().sum(0)
    if mn:  # stored transposed: [K, M] / [K, N]
        A_dev = A.transpose(1, 2).contiguous()
        B_dev = B.transpose(1, 2).contiguous()
        lda, ldb = M, N
    else:
        A_dev, B_dev = A.contiguous(), B.contiguous()
        lda, ldb = K, K
    out = torch.full((planes, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    out2 = torch.full((planes, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    outf = torch.zeros(M, N, device=dev)
    ref = Af.double() @ Bf.double().t()
    if epi in (1, 2, 3):
        ref = ref + bias.double()
    if epi in (3, 5):
        ref = ref + auxf.double()
    ref2 = None
    if epi == 2:  # out = gelu'(pre), out2 = gelu(pre)
        ref2 = torch.nn.functional.gelu(ref)
        cdf = 0.5 * (1 + torch.erf(ref / 2**0.5))
        pdf = torch.exp(-0.5 * ref * ref) / (2 * 3.141592653589793) ** 0.5
        ref = cdf + ref * pdf
    if epi == 4:  # out = acc * aux
        ref = ref * auxf.double()
    stream = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    st = lib.simx_gemm(M, N, K, A_dev.data_ptr(), lda, A_dev.stride(0), mn, B_dev.data_ptr(), ldb, B_dev.stride(0), mn,
                       planes, epi, bias.data_ptr(), aux.data_ptr(), N, aux.stride(0), out.data_ptr(), out2.data_ptr(),
                       N, out.stride(0), outf.data_ptr(), N, ksplit, stream)
    if st != 0:
        return {"case": name, "ok": False, "error": lib.simx_last_error().decode()}
    torch.cuda.synchronize()
    got = outf.double() if epi == 6 else out.float().sum(0).double()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    res = {"case": name, "rel_err": err, "ms_first": (time.time() - t0) * 1e3}
    tol = 1e-5 if planes == 3 else (2e-5 if epi == 6 else (1e-2 if epi == 2 else 6e-3))
    if planes == 3 and epi != 6:
        tol = 1e-5
    if ref2 is not None:
        got2 = out2.float().sum(0).double()
        err2 = (got2 - ref2).abs().max().item() / ref2.abs().max().item()
        res["rel_err2"] = err2
        err = max(err, err2)
    res["tol"] = tol
    res["ok"] = bool(err < tol) and not bool(torch.isnan(got).any())
    # timing (warm)
    if res["ok"]:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            lib.simx_gemm(M, N, K, A_dev.data_ptr(), lda, A_dev.stride(0), mn, B_dev.data_ptr(), ldb, B_dev.stride(0),
                          mn, planes, epi, bias.data_ptr(), aux.data_ptr(), N, aux.stride(0), out.data_ptr(),
                          out2.data_ptr(), N, out.stride(0), outf.data_ptr(), N, ksplit, stream)
        evs[0].record()
        for _ in range(10):
            lib.simx_gemm(M, N, K, A_dev.data_ptr(), lda, A_dev.stride(0), mn, B_dev.data_ptr(), ldb, B_dev.stride(0),
                          mn, planes, epi, bias.data_ptr(), aux.data_ptr(), N, aux.stride(0), out.data_ptr(),
                          out2.data_ptr(), N, out.stride(0), outf.data_ptr(), N, ksplit, stream)
        evs[1].record()
        torch.cuda.synchronize()
        ms = evs[0].elapsed_time(evs[1]) / 10
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K * (6 if planes == 3 else 1) / ms / 1e9
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "check_gemm.json"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    if args.case >= 0:
        print("RESULT " + json.dumps(run_case(args.case)))
        return 0
    results = []
    for i, c in enumerate(CASES):
        if args.only and args.only not in c[0]:
            continue
        try:
            proc = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True, text=True,
                                  timeout=120)
            line = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                r = json.loads(line[-1][7:])
            else:
                r = {"case": c[0], "ok": False, "error": (proc.stderr or proc.stdout)[-1500:]}
        except subprocess.TimeoutExpired:
            r = {"case": c[0], "ok": False, "error": "TIMEOUT (hang)"}
        print(json.dumps(r), flush=True)
        results.append(r)
        if r.get("error", "").startswith("TIMEOUT"):
            # a hung kernel may have wedged the context; stop rather than pile up strikes
            break
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)
    bad = [r for r in results if not r.get("ok")]
    print(f"{len(results) - len(bad)}/{len(results)} ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
