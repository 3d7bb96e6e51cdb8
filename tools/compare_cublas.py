"""Sustained bf16 GEMM rate of libsimx_b200's tcgen05 kernel (plain epilogue) against cuBLAS (torch.nn.functional.linear)
at the encoder's own shapes, each timed over >= 1.5 s of back-to-back launches so both run at the power-capped clock."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from simxns_b200 import _lib  # noqa: E402

T = 262144
SHAPES = [("qkv", T, 2304, 768), ("attn_out", T, 768, 768), ("ffn_up", T, 3072, 768), ("ffn_down", T, 768, 3072),
          ("square", 8192, 8192, 8192)]


def sustained(fn, seconds=1.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for name, M, N, K in SHAPES:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        B = (0.02 * torch.randn(N, K, device="cuda")).to(torch.bfloat16)
        D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

        def ours():
            _lib.check(lib.simx_gemm(M, N, K, A.data_ptr(), K, 0, 0, B.data_ptr(), K, 0, 0, 1, _lib.EPI_PLAIN, None, None, 0, 0,
                                     D.data_ptr(), None, N, 0, None, 0, 1, st), "simx_gemm")

        def cublas():
            torch.nn.functional.linear(A, B, out=None) if False else torch.matmul(A, B.t(), out=D)

        r = dict(shape=name, M=M, N=N, K=K)
        for tag, fn in (("cublas", cublas), ("simx", ours), ("cublas2", cublas), ("simx2", ours)):
            ms = sustained(fn)
            r[tag + "_tflops"] = 2.0 * M * N * K / ms / 1e9
        out.append(r)
        print(json.dumps(r), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/r01_vs_cublas.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
