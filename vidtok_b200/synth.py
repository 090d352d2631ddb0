"""Deterministic synthetic weights / inputs (no checkpoints are available offline).

The recipe depends only on (sorted key, shape, seed) -- not on module construction order -- so the
reference model, the oracle and the CUDA path can all be loaded with bit-identical tensors.
Every tensor is non-trivial: in particular the temporal ResBlocks' conv2, which the reference
zero-initialises (vidtok/modules/model_3dcausal.py:460-462), gets real values so the temporal path is
actually exercised.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Sequence

import torch


def synth_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for key in sorted(shapes.keys()):
        shape = tuple(int(s) for s in shapes[key])
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        leaf = key.rsplit(".", 1)[-1]
        is_norm = ".norm" in key  # norm1/norm2/norm_out/attn norm; keys end in .norm.weight or .weight
        if leaf == "mix_factor":
            t = 1.0 + 0.5 * r
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = r * (1.0 / math.sqrt(fan_in))
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * r
        elif is_norm and leaf == "bias":
            t = 0.1 * r
        else:
            t = 0.05 * r
        out[key] = t.contiguous()
    return out


def synth_clip(batch: int, frames: int, height: int, width: int, seed: int = 1234, channels: int = 3) -> torch.Tensor:
    """Uniform [-1,1] video tensor [B,3,T,H,W] (the README idiom, reference README.md:336)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand((batch, channels, frames, height, width), generator=g, dtype=torch.float32) * 2 - 1


def synth_noise(shape, seed: int = 4321) -> torch.Tensor:
    """Stand-in for the reference's torch.randn(mean.shape) on the CPU generator
    (vidtok/modules/distributions.py:17) with an explicit generator."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32)


def weights_fingerprint(sd: Mapping[str, torch.Tensor]) -> float:
    tot = 0.0
    for k in sorted(sd.keys()):
        tot += float(sd[k].double().abs().sum())
    return tot
