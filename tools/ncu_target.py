"""Minimal target for Nsight Compute: `forwards` full encode->decode passes of a bench workload.
Usage: python tools/ncu_target.py [B] [forwards] [bf16|exact|mixed|fma] [kl488|fsq488|v11long|kl41616]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vidtok_b200.compat_util import instantiate_from_config  # noqa: E402
from vidtok_b200.synth import synth_clip, synth_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
c = bench.CONFIGS[sys.argv[4] if len(sys.argv) > 4 else "kl488"]
model = instantiate_from_config(bench.model_cfg(c))
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0))
model = model.cuda().eval()
model.precision = prec
if c["tiling"]:
    model.use_tiling = True
    model.t_chunk_enc, model.t_chunk_dec, model.use_overlap = c["tiling"]
x = synth_clip(B, c["T"], c["H"], c["W"]).cuda()
with torch.no_grad():
    for _ in range(n):
        model(x)
torch.cuda.synchronize()
print("ncu_target done")
