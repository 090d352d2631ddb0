"""The drop-in boundary of the reference: `instantiate_from_config` (vidtok/modules/util.py:69-86) resolves a YAML
`target:` string to a class and calls it with `params`.  Same behaviour, same error for a missing `target`."""
from __future__ import annotations

import importlib

import torch


def get_obj_from_str(string: str, reload: bool = False, invalidate_cache: bool = True):
    module_name, cls_name = string.rsplit(".", 1)
    if invalidate_cache:
        importlib.invalidate_caches()
    module = importlib.import_module(module_name, package=None)
    if reload:
        module = importlib.reload(module)
    return getattr(module, cls_name)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    params = config.get("params", dict())
    return get_obj_from_str(config["target"])(**params)


def print0(*args, **kwargs):
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            return
    except Exception:
        pass
    print(*args, **kwargs)


def compute_psnr(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Per-frame PSNR averaged over frames, inputs in [0,1] (vidtok/modules/util.py:146-155)."""
    if x.dim() == 5:
        assert y.dim() == 5
        x = x.permute(0, 2, 1, 3, 4).reshape(-1, x.shape[1], x.shape[3], x.shape[4])
        y = y.permute(0, 2, 1, 3, 4).reshape(-1, y.shape[1], y.shape[3], y.shape[4])
    mse = ((x - y) ** 2).mean(dim=[1, 2, 3])
    return (-10.0 * torch.log10(mse + 1e-8)).mean(dim=0)
