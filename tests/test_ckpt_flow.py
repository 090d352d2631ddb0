"""The caller-side flow of the reference's inference scripts, replayed without omegaconf/decord:
YAML -> instantiate_from_config(config.model) with ckpt_path / ignore_keys / verbose set by the caller
(scripts/inference_evaluate.py:26-32) -> init_from_ckpt for `.ckpt` and `.safetensors` (autoencoder.py:146-176)
-> model.to(device).eval() -> model(input) (scripts/inference_evaluate.py:141,173)."""
import copy
import os

import pytest
import torch
import yaml

from conftest import GOLDEN_DIR, load_golden, synth_inputs, synth_weights


def load_model_from_config(cfg_path, ckpt, ignore_keys=(), verbose=False):
    cfg = yaml.safe_load(open(cfg_path))
    m = cfg["model"]
    # OmegaConf would resolve ${model.params.encoder_config.params}
    if isinstance(m["params"]["decoder_config"]["params"], str):
        assert m["params"]["decoder_config"]["params"] == "${model.params.encoder_config.params}"
        m["params"]["decoder_config"]["params"] = copy.deepcopy(m["params"]["encoder_config"]["params"])
    m["params"]["ckpt_path"] = ckpt
    m["params"]["ignore_keys"] = list(ignore_keys)
    m["params"]["verbose"] = verbose
    from vidtok_b200.compat_util import instantiate_from_config
    return instantiate_from_config(m)


@pytest.fixture(scope="module")
def ckpts(tmp_path_factory):
    d, meta = load_golden("tiny_kl_v10")
    sd = synth_weights(meta, d)
    full = dict(sd)
    full["loss.logvar"] = torch.zeros(())                       # released checkpoints carry the loss module
    full["loss.discriminator.main.0.weight"] = torch.zeros(4, 3, 4, 4)
    root = tmp_path_factory.mktemp("ckpt")
    p1 = str(root / "model.ckpt")
    torch.save({"state_dict": full, "global_step": 123}, p1)
    p2 = str(root / "bare.ckpt")
    torch.save(full, p2)
    p3 = str(root / "model.safetensors")
    from safetensors.torch import save_file
    save_file({k: v.contiguous() for k, v in full.items()}, p3)
    return sd, (p1, p2, p3)


def test_init_from_ckpt_variants(ckpts):
    sd, paths = ckpts
    cfg = os.path.join(GOLDEN_DIR, "cfg_kl_488_4chn_model.yaml")
    for p in paths:
        model = load_model_from_config(cfg, p)
        got = model.state_dict()
        assert set(got) == set(sd)
        assert all(torch.equal(got[k], sd[k]) for k in sd)
    # ignore_keys are regexes matched with re.match (autoencoder.py:156-162): dropped keys keep their initial values
    model = load_model_from_config(cfg, paths[0], ignore_keys=[r"decoder\.conv_out"])
    got = model.state_dict()
    assert not torch.equal(got["decoder.conv_out.conv.weight"], sd["decoder.conv_out.conv.weight"])
    assert torch.equal(got["decoder.conv_in.conv.weight"], sd["decoder.conv_in.conv.weight"])
    with pytest.raises(NotImplementedError):
        load_model_from_config(cfg, "weights.bin")


@pytest.mark.gpu
def test_script_flow_matches_reference_fixture(ckpts):
    sd, paths = ckpts
    d, meta = load_golden("tiny_kl_v10")
    x = synth_inputs(meta, d)
    model = load_model_from_config(os.path.join(GOLDEN_DIR, "cfg_kl_488_4chn_model.yaml"), paths[2])
    device = torch.device("cuda")
    model.to(device).eval()
    assert model.is_causal and model.encoder.time_downsample_factor == 4 and not hasattr(model, "use_tiling")
    with torch.no_grad():
        torch.manual_seed(meta["noise_seed"])
        _, output, reg_log = model(x.to(device))
    output = output.clamp(-1, 1)
    assert float((output.cpu() - torch.from_numpy(d["dec"]).clamp(-1, 1)).abs().max()) <= 1e-3
    assert "kl_loss" in reg_log
