"""Launch one GEMM configuration a few times (for `ncu --set full -k regex:gemm_kernel`).
    python tools/profile_gemm.py ffn1_fwd|qkv_fwd|ffn2_fwd|ffn1_wgrad [--tokens 65536] [--iters 3]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from simxns_b200 import _lib  # noqa: E402

H, I = 768, 3072
CASES = {  # name: (M, N, K, mn_major, epilogue)   M == 'T' means the token count
    "qkv_fwd": ("T", 3 * H, H, 0, _lib.EPI_BIAS),
    "attnout_fwd": ("T", H, H, 0, _lib.EPI_BIAS_ADD),
    "ffn1_fwd": ("T", I, H, 0, _lib.EPI_BIAS_GELU),
    "ffn2_fwd": ("T", H, I, 0, _lib.EPI_BIAS_ADD),
    "ffn2_dgrad": ("T", I, H, 0, _lib.EPI_MUL),
    "ffn1_dgrad": ("T", H, I, 0, _lib.EPI_ADD),
    "ffn1_wgrad": (I, H, "T", 1, _lib.EPI_ATOMIC_F32),
    "attnout_dgrad": ("T", H, H, 0, _lib.EPI_PLAIN),
    "qkv_dgrad": ("T", H, 3 * H, 0, _lib.EPI_ADD),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--tokens", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--epi", type=int, default=-1, help="override the epilogue (SIMX_EPI_*) of the case's shape")
    a = ap.parse_args()
    M, N, K, mn, epi = [a.tokens if v == "T" else v for v in CASES[a.case]]
    if a.epi >= 0:
        epi = a.epi
    lib = _lib.lib()
    dev = "cuda"
    A = torch.randn(K, M, device=dev).to(torch.bfloat16) if mn else torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(K, N, device=dev).to(torch.bfloat16) if mn else (0.02 * torch.randn(N, K, device=dev)).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    aux = torch.randn(M, N, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    outf = torch.zeros(M, N, device=dev) if epi == _lib.EPI_ATOMIC_F32 else torch.zeros(N, device=dev)  # EPI_MUL: column sums
    lda, ldb = (M, N) if mn else (K, K)
    ks = 1
    if mn:
        tiles = ((M + 127) // 128) * ((N + 255) // 256)
        ks = max(1, min((K // 64) // 4, (3 * 148 + tiles - 1) // tiles))
    st = torch.cuda.current_stream().cuda_stream
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(a.iters + 1):
        if i == 1:
            evs[0].record()
        _lib.check(lib.simx_gemm(M, N, K, A.data_ptr(), lda, 0, mn, B.data_ptr(), ldb, 0, mn, 1, epi, bias.data_ptr(),
                                 aux.data_ptr(), N, 0, out.data_ptr(), out2.data_ptr(), N, 0, outf.data_ptr(), N, ks, st),
                   "simx_gemm")
    evs[1].record()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[1]) / a.iters
    print(f"{a.case} epi={epi}: M={M} N={N} K={K} {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
