"""Kernel-time attribution of the tiled long-video path (config 4) -- kernels vs host overhead."""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_full import build, make_cfg  # noqa: E402
from vidtok_b200 import _native as N  # noqa: E402
from vidtok_b200.synth import synth_clip  # noqa: E402

cfg = make_cfg(version="v1_1", z=16, interp="trilinear")
model, sd = build(cfg)
model.use_tiling, model.t_chunk_enc, model.t_chunk_dec, model.use_overlap = True, 16, 4, True
model.precision = "bf16"
x = synth_clip(1, 129, 256, 256, seed=7).cuda()
lib = N.lib()
with torch.no_grad():
    model(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model(x)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    lib.vt_profile_start()
    model(x)
    buf = ctypes.create_string_buffer(1 << 20)
    lib.vt_profile_stop(buf, len(buf))
prof = json.loads(buf.value.decode())
tot = sum(v["ms"] for v in prof.values())
print(f"wall {wall * 1e3:.1f} ms; sum of kernel ms {tot:.1f} over {sum(v['launches'] for v in prof.values())} launches")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]:
    print(f"{v['ms']:9.3f} ms n={v['launches']:5d} {k}")
