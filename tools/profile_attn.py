"""Standalone attention forward/backward launches (for ncu and timing): python tools/profile_attn.py [--B 2048] [--S 128]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from simxns_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2048)
    ap.add_argument("--S", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    H, heads = 768, 12
    lib = _lib.lib()
    T = a.B * a.S
    qkv = (0.5 * torch.randn(T, 3 * H, device="cuda")).to(torch.bfloat16)
    ctx = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
    dctx = torch.randn(T, H, device="cuda").to(torch.bfloat16)
    dqkv = torch.empty(T, 3 * H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(a.B, heads, a.S, device="cuda")
    delta = torch.empty(a.B, heads, a.S, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def fwd():
        _lib.check(lib.simx_attention_forward(qkv.data_ptr(), ctx.data_ptr(), None, lse.data_ptr(), a.B, a.S, H, heads, 1, 0,
                                              0, st), "fwd")

    def bwd():
        _lib.check(lib.simx_attention_backward(qkv.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), dqkv.data_ptr(), None,
                                               lse.data_ptr(), delta.data_ptr(), a.B, a.S, H, heads, 1, 0, 0, None, st), "bwd")

    import ctypes
    tbuf = torch.zeros(8, dtype=torch.int64, device="cuda")
    lib.simx_debug_attention_timing.argtypes = [ctypes.c_void_p]
    lib.simx_debug_attention_timing(tbuf.data_ptr())
    fwd(); bwd(); bwd()
    torch.cuda.synchronize()
    t = tbuf.cpu().tolist()
    print("bwd CTA phase cycles:", [t[i + 1] - t[i] for i in range(7)], "total", t[7] - t[0])
    lib.simx_debug_attention_timing(None)
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"attn {name}: B={a.B} S={a.S}: {e0.elapsed_time(e1) / a.iters:.3f} ms")


if __name__ == "__main__":
    main()
