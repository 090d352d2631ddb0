"""Minimal target for Nsight Compute: `forwards` full encode->decode passes of the bench workload
(vidtok_kl_causal_488_4chn, bf16, B clips of 17x256x256).  Usage: python tools/ncu_target.py [B] [forwards]"""
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
model = instantiate_from_config(bench.model_cfg())
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0))
model = model.cuda().eval()
model.precision = "bf16"
x = synth_clip(B, 17, 256, 256).cuda()
with torch.no_grad():
    for _ in range(n):
        model(x)
torch.cuda.synchronize()
print("ncu_target done")
