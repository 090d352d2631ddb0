"""CPU-only checks of bench.py's host logic: the four configurations against BASELINE.json / SURVEY §6, the reference arm's
model construction without the CUDA library, and the JSON contract keys of the reference arm on a tiny clip."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_configs_match_the_baseline_table():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "kl_causal_488" in bench.METRIC and "kl_causal_488" in base["metric"]
    c = bench.CONFIGS
    assert set(c) == {"kl488", "fsq488", "v11long", "kl41616"}
    # the headline stays the default (BENCH / SCALE records of the driver)
    ap_default = [a for a in open(os.path.join(ROOT, "bench.py")).read().splitlines() if '"--config"' in a][0]
    assert 'default="kl488"' in ap_default
    # algorithmic FLOPs per clip (SURVEY §6 / BASELINE.md §2)
    assert c["kl488"]["flops"] == pytest.approx(20.691e12) and c["fsq488"]["flops"] == pytest.approx(20.690e12)
    assert c["v11long"]["flops"] == pytest.approx(160.38e12) and c["kl41616"]["flops"] == pytest.approx(85.627e12)
    assert c["fsq488"]["precision"] == "mixed" and c["kl488"]["precision"] == "bf16"
    assert c["v11long"]["tiling"] == (16, 4, True) and c["v11long"]["T"] == 129


@pytest.mark.parametrize("name,tensors", [("kl488", 416), ("fsq488", 416), ("kl41616", None), ("v11long", None)])
def test_reference_arm_builds_its_model_from_the_shape_table(name, tensors):
    """`--impl reference` must not touch the CUDA library: the weight manifest comes from the oracle's parameter table
    (416 tensors / 157.4 M parameters for the 488 models)."""
    from oracle.vidtok_oracle import cfg_from_model_yaml, reference_param_shapes
    shapes = reference_param_shapes(cfg_from_model_yaml(bench.model_cfg(bench.CONFIGS[name])))
    n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    if tensors is not None:
        assert len(shapes) == tensors
        assert 157.0e6 < n < 157.8e6
    assert "encoder.conv_in.conv.weight" in shapes and "decoder.conv_out.conv.weight" in shapes


def test_importing_bench_does_not_load_the_cuda_library():
    import subprocess
    code = "import sys; sys.path.insert(0, %r); import bench; import ctypes; print(any('libvidtok_b200' in l for l in open('/proc/self/maps')))" % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip().endswith("False")


def test_cpu_sample_scaling_is_in_full_size_frames():
    c = bench.CONFIGS["kl488"]
    assert bench.cpu_units_scale(c, 17, 256) == pytest.approx(17.0)
    assert bench.cpu_units_scale(c, 17, 128) == pytest.approx(17.0 / 4)
    v = bench.CONFIGS["v11long"]
    assert bench.cpu_units_scale(v, 33, 256) == pytest.approx(33.0)
