"""Target for Nsight Compute: the fused temporal ResBlock alone at the bench geometry (B x 20 x 256 x 256 x 128, bf16).
Usage: python tools/run_tblock.py [B] [launches] [ln_out]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vidtok_b200 import _native as N  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ln = int(sys.argv[3]) if len(sys.argv) > 3 else 1
T, H, W, Cc = 20, 256, 256, 128
g = torch.Generator(device="cuda").manual_seed(0)
n1 = torch.randn((B, T, H, W, Cc), device="cuda", generator=g).to(torch.bfloat16)
x = torch.randn((B, T, H, W, Cc), device="cuda", generator=g).to(torch.bfloat16)
w1 = torch.randn((Cc, Cc, 3), device="cuda", generator=g) / (3 * Cc) ** 0.5
w2 = torch.randn((Cc, Cc, 3), device="cuda", generator=g) / (3 * Cc) ** 0.5
v = [torch.randn(Cc, device="cuda", generator=g) for _ in range(6)]
out, out2 = torch.empty_like(x), torch.empty_like(x)
p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
for i in range(n):
    ev[i].record()
    N.check(N.lib().vt_op_tblock(p(n1), p(x), p(w1), p(v[0]), p(v[1]), p(v[2]), p(w2), p(v[3]), p(v[4]), p(v[5]), 1, p(out),
                                 p(out2) if ln else None, B, T, H, W, Cc, None))
ev[n].record()
torch.cuda.synchronize()
print("tblock ms per call (incl. weight repack + sync):", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(n)])
