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


def compute_ssim(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Mean SSIM over frames and channels, inputs in [0,1] (vidtok/modules/util.py:157-178: 11-tap Gaussian window with
    sigma 1.5, k1 = 0.01, k2 = 0.03, frames average-pooled by round(min(H, W) / 256) first).  Works on CPU and CUDA tensors.
    The window is applied as two 1-D passes (the 2-D Gaussian is separable) over the five moment maps at once."""
    import torch.nn.functional as F

    if x.dim() == 5:
        assert y.dim() == 5
        x = x.permute(0, 2, 1, 3, 4).reshape(-1, x.shape[1], x.shape[3], x.shape[4])
        y = y.permute(0, 2, 1, 3, 4).reshape(-1, y.shape[1], y.shape[3], y.shape[4])
    f = max(1, round(min(x.shape[-2:]) / 256))
    if f > 1:
        x, y = F.avg_pool2d(x, kernel_size=f), F.avg_pool2d(y, kernel_size=f)
    size, sigma, c1, c2 = 11, 1.5, 0.01 ** 2, 0.03 ** 2
    if x.shape[-1] < size or x.shape[-2] < size:
        raise ValueError(f"Kernel size can't be greater than actual input size. Input size: {x.size()}. Kernel size: {size}")
    t = torch.arange(size, dtype=x.dtype, device=x.device) - (size - 1) / 2.0
    g = torch.exp(-(t ** 2) / (2 * sigma ** 2))
    g = g / g.sum()
    n, c = x.shape[:2]
    maps = torch.cat([x, y, x * x, y * y, x * y], dim=0)                      # [5n, c, H, W]
    maps = F.conv2d(maps, g.view(1, 1, 1, size).repeat(c, 1, 1, 1), groups=c)
    maps = F.conv2d(maps, g.view(1, 1, size, 1).repeat(c, 1, 1, 1), groups=c)
    mu_x, mu_y, e_xx, e_yy, e_xy = maps.split(n, dim=0)
    s_xx, s_yy, s_xy = e_xx - mu_x * mu_x, e_yy - mu_y * mu_y, e_xy - mu_x * mu_y
    cs = (2.0 * s_xy + c2) / (s_xx + s_yy + c2)
    ss = (2.0 * mu_x * mu_y + c1) / (mu_x * mu_x + mu_y * mu_y + c1) * cs
    return ss.mean(dim=(-1, -2)).mean(dim=1).mean(dim=0)


# small helpers other reference modules import from vidtok.modules.util (lpips.py:11, logger.py:20, losses.py:11, data/*.py)
def exists(x):
    return x is not None


def default(val, d):
    from inspect import isfunction
    if val is not None:
        return val
    return d() if isfunction(d) else d


def isheatmap(x):
    return isinstance(x, torch.Tensor) and x.ndim == 2


def seed_anything(seed: int):
    import os
    import random

    import numpy as np
    os.environ["PYTHONHASHSEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_valid_dirs(*dirs):
    import os
    for d in dirs:
        if d is not None and os.path.isdir(d):
            return d
    return None


def get_valid_paths(*paths):
    import os
    for p in paths:
        if p is not None and os.path.isfile(p):
            return p
    return None
