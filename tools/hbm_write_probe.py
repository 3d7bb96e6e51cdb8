"""Pure-write, pure-read and copy bandwidth of this GPU (torch fill_ / sum / copy_ over 4 GiB, CUDA events, best of 10):
the denominators for kernels whose traffic is mostly stores (the FFN-up GEMM writes two [T, 3072] bf16 tensors)."""
import json

import torch

n = 2 * 1024 ** 3  # bf16 elements = 4 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")


def best(fn, bytes_moved, reps=10):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return bytes_moved / (min(t) * 1e-3) / 1e9


res = {
    "write_gbs (fill_)": best(lambda: a.fill_(1.0), 2 * n),
    "read_gbs (sum)": best(lambda: a.view(torch.int16).sum(dtype=torch.int64), 2 * n),
    "copy_gbs (read + write)": best(lambda: b.copy_(a), 4 * n),
}
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
