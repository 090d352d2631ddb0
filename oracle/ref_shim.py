"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (vidtok_b200/).

Import helper for the UNMODIFIED reference at /root/reference (only present in the authoring
container, never on the GPU box).  The reference needs `lightning` and `omegaconf`, which are not
installed here; the only symbols on the inference path are
`lightning.pytorch.LightningModule` (vidtok/models/autoencoder.py:9,18),
`lightning.pytorch.utilities.rank_zero.rank_zero_only` (vidtok/modules/util.py:13) and
`omegaconf.ListConfig` (vidtok/models/autoencoder.py:5).  We stub exactly those.

Used by oracle/make_golden.py to generate tests/golden/*.npz and by the (container-only) test that
pins oracle/vidtok_oracle.py against the real reference.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VIDTOK_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vidtok", "modules"))


def _install_stubs():
    import torch

    if "lightning" not in sys.modules:
        try:
            import lightning  # noqa: F401
        except Exception:
            lightning = types.ModuleType("lightning")
            pl = types.ModuleType("lightning.pytorch")

            class LightningModule(torch.nn.Module):
                global_step = 0

            pl.LightningModule = LightningModule
            utilities = types.ModuleType("lightning.pytorch.utilities")
            rank_zero = types.ModuleType("lightning.pytorch.utilities.rank_zero")
            rank_zero.rank_zero_only = lambda f: f
            utilities.rank_zero = rank_zero
            utilities.rank_zero_only = rank_zero.rank_zero_only
            pl.utilities = utilities
            lightning.pytorch = pl
            sys.modules["lightning"] = lightning
            sys.modules["lightning.pytorch"] = pl
            sys.modules["lightning.pytorch.utilities"] = utilities
            sys.modules["lightning.pytorch.utilities.rank_zero"] = rank_zero
    if "omegaconf" not in sys.modules:
        try:
            import omegaconf  # noqa: F401
        except Exception:
            omegaconf = types.ModuleType("omegaconf")
            omegaconf.ListConfig = list
            sys.modules["omegaconf"] = omegaconf


def import_reference():
    """Returns the reference `vidtok` package (imported from REFERENCE_ROOT, unmodified).

    This repo ships a regular `vidtok` package with the same import paths (the drop-in shims); a regular package anywhere
    on sys.path beats the reference's namespace package, so every sys.path entry that holds a `vidtok/__init__.py` is
    dropped for this process and the already-imported shim modules are forgotten before the reference is imported."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    for k in [k for k in sys.modules if k == "vidtok" or k.startswith("vidtok.")]:
        del sys.modules[k]
    ref = os.path.abspath(REFERENCE_ROOT)
    keep = []
    for p in sys.path:
        q = os.path.abspath(p or ".")
        if q != ref and os.path.isfile(os.path.join(q, "vidtok", "__init__.py")):
            continue   # the shim package's parent directory (usually the repo root / the script's cwd)
        if q != ref:
            keep.append(p)
    sys.path[:] = [ref] + keep
    import importlib
    importlib.invalidate_caches()
    import vidtok  # noqa: F401
    import vidtok.models.autoencoder  # noqa: F401
    import vidtok.models.autoencoder_v1_1  # noqa: F401

    assert all(os.path.abspath(p).startswith(ref) for p in vidtok.__path__), list(vidtok.__path__)
    return vidtok


def build_reference_model(model_cfg: dict):
    """model_cfg = the `model:` section of a reference YAML (dict), with the
    `${model.params.encoder_config.params}` interpolation resolved by hand and the loss replaced by
    torch.nn.Identity (skips the LPIPS/VGG download, vidtok/modules/lpips.py:55)."""
    import copy

    import_reference()
    from vidtok.modules.util import instantiate_from_config

    cfg = copy.deepcopy(model_cfg)
    p = cfg["params"]
    if isinstance(p["decoder_config"].get("params"), str):
        p["decoder_config"]["params"] = copy.deepcopy(p["encoder_config"]["params"])
    p["loss_config"] = {"target": "torch.nn.Identity"}
    p.pop("ckpt_path", None)
    model = instantiate_from_config(cfg)
    return model.eval()
