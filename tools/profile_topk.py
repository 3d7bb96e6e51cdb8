"""clock64 phase breakdown of one top-k select CTA (debug hook simx_debug_topk_timing) + kernel times."""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from simxns_b200 import _lib
from simxns_b200.search import FlatIPIndex

lib = _lib.lib()
N, H, nq = 300000, 768, 2048
g = torch.Generator(device="cuda").manual_seed(0)
base = torch.randn(H, device="cuda", generator=g)
index = FlatIPIndex(H, capacity=N)
for b in range(0, N, 100000):
    index.add(0.5 * torch.randn(min(100000, N - b), H, device="cuda", generator=g) + base)
q = 0.5 * torch.randn(nq, H, device="cuda", generator=g) + base
tbuf = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.simx_debug_topk_timing.argtypes = [ctypes.c_void_p]
for k in (200, 1000):
    index.search_local(q, k)
    lib.simx_debug_topk_timing(tbuf.data_ptr())
    index.search_local(q, k)
    torch.cuda.synchronize()
    lib.simx_debug_topk_timing(None)
    t = tbuf.cpu().tolist()
    print(f"k={k}: cycles minmax+clear {t[1]-t[0]} hist {t[2]-t[1]} find_bin {t[3]-t[2]} flags+scan {t[7]-t[3]} "
          f"writes {t[4]-t[7]} rank+stage {t[5]-t[4]} store {t[6]-t[5]} | total after load {t[6]-t[0]}")
