"""BASELINE config 4: vidtok_kl_causal_488_16chn v1.1, one 129x256x256 video, tiled (t_chunk_enc=16, overlap) -- frames/s."""
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
with torch.no_grad():
    for _ in range(2):
        model(x)
    torch.cuda.synchronize()
    N.lib().vt_launch_count(1)
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        z, dec, _ = model(x)
    torch.cuda.synchronize()
    el = (time.perf_counter() - t0) / n
print(f"config4 tiled 129x256x256 bf16: {el * 1e3:.1f} ms per video, {129 / el:.1f} frames/s, {N.lib().vt_launch_count(0) // n} launches per video; "
      f"160.38 TFLOP/video -> {160.38 / el:.0f} TFLOP/s")
