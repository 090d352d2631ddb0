"""The oracle (oracle/vidtok_oracle.py) against the golden fixtures that oracle/make_golden.py produced by running the
UNMODIFIED reference.  Runs anywhere (no /root/reference, no GPU)."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden, synth_inputs, synth_weights


def run_oracle(meta, sd, x):
    from oracle.vidtok_oracle import OracleModel, cfg_from_model_yaml
    om = OracleModel(cfg_from_model_yaml(meta["model"]), sd)
    if meta["tiling_chunk"]:
        om.use_tiling, om.t_chunk_enc, om.use_overlap = True, meta["tiling_chunk"], True
        om.t_chunk_dec = om.t_chunk_enc // 4
    torch.manual_seed(meta["noise_seed"])
    z, log, h = om.encode(x, return_pre=True)
    dec = om.decode(z)
    if dec.shape[2] != x.shape[2] and om.cfg.version == "v1_1":
        dec = dec[:, :, -x.shape[2]:]
    return z, dec, log, h


def check_fsq_indices(idx, idx_ref, pre_round):
    """FSQ equality: exact, or mismatching only where bound(z) sits within 1e-4 of a rounding tie."""
    idx, idx_ref = np.asarray(idx), np.asarray(idx_ref)
    bad = idx != idx_ref
    if not bad.any():
        return 0
    frac = np.abs(pre_round - np.floor(pre_round) - 0.5)  # distance to the nearest .5
    near_tie = (frac < 1e-4).any(axis=-1)
    assert not (bad & ~near_tie).any(), f"{int((bad & ~near_tie).sum())} FSQ index mismatches away from rounding ties"
    return int(bad.sum())


@pytest.mark.parametrize("case", [c for c in golden_cases() if not c.startswith("cfg1")])
def test_oracle_reproduces_reference_small(case):
    d, meta = load_golden(case)
    sd, x = synth_weights(meta, d), synth_inputs(meta, d)
    z, dec, log, h = run_oracle(meta, sd, x)
    assert float((z - torch.from_numpy(d["z"])).abs().max()) <= 1e-4
    assert dec.shape == tuple(d["dec"].shape)
    assert float((dec - torch.from_numpy(d["dec"])).abs().max()) <= 2e-4
    if "indices" in d:
        assert log["indices"].dtype == torch.int32
        check_fsq_indices(log["indices"].numpy(), d["indices"], log["pre_round"].numpy())
    else:
        assert abs(float(log["kl_loss"]) - float(d["kl_loss"])) <= 1e-4 * abs(float(d["kl_loss"]))


@pytest.mark.parametrize("case", golden_cases("cfg1"))
def test_oracle_reproduces_reference_config1(case):
    """BASELINE.json configs[0]: 1x3x17x128x128 through the full-size 488 model."""
    d, meta = load_golden(case)
    sd, x = synth_weights(meta, d), synth_inputs(meta, d)
    z, dec, log, h = run_oracle(meta, sd, x)
    assert float((z - torch.from_numpy(d["z"])).abs().max()) <= 1e-4
    sel = [int(i) for i in d["dec_frames"]]
    assert float((dec[:, :, sel] - torch.from_numpy(d["dec_sel"])).abs().max()) <= 2e-4
    assert np.allclose(dec.double().mean(dim=(0, 1, 3, 4)).numpy(), d["dec_frame_mean"], atol=1e-5)
    if "indices" in d:
        check_fsq_indices(log["indices"].numpy(), d["indices"], log["pre_round"].numpy())


def test_fsq_known_answers():
    """Properties of the FSQ arithmetic that do not depend on torch numerics (regularizers.py:114,153-198)."""
    from oracle.vidtok_oracle import fsq_constants, fsq_indices_to_codes, fsq_regularize
    lv, basis, half_l, offset, shift = fsq_constants((8, 8, 8, 8, 8))
    assert basis.tolist() == [1, 8, 64, 512, 4096]
    assert abs(float(half_l[0]) - 3.5035) < 1e-6 and float(offset[0]) == 0.5
    idx = torch.arange(32768, dtype=torch.int32).reshape(1, 8, 64, 64)
    codes = fsq_indices_to_codes(idx, (8, 8, 8, 8, 8))
    assert set(np.unique(codes.numpy()).tolist()) == {-1.0, -0.75, -0.5, -0.25, 0.0, 0.25, 0.5, 0.75}
    # codes_to_indices(indices_to_codes(i)) == i : feed values that quantise back onto the same codes
    z = torch.atanh((codes * 4 + 0.5) / 3.5035) - shift.view(1, 5, 1, 1, 1)
    _, log = fsq_regularize(z, (8, 8, 8, 8, 8))
    assert torch.equal(log["indices"], idx)
    lv2, basis2, *_ = fsq_constants((8, 8, 8, 5, 5, 5))
    assert basis2.tolist() == [1, 8, 64, 512, 2560, 12800]


def test_chunk_schedule_and_frame_algebra():
    from oracle.vidtok_oracle import build_chunk_start_end
    assert build_chunk_start_end(129, 16)[:3] == [[0, 1], [1, 17], [17, 33]] and len(build_chunk_start_end(129, 16)) == 9
    assert build_chunk_start_end(33, 4) == [[0, 1], [1, 5], [5, 9], [9, 13], [13, 17], [17, 21], [21, 25], [25, 29], [29, 33]]
    assert build_chunk_start_end(1, 16) == [[0, 1]]
    from vidtok_b200.compat_util import instantiate_from_config
    d, meta = load_golden("tiny_kl_v11_tiled")
    from conftest import resolved_model_cfg
    model = instantiate_from_config(resolved_model_cfg(meta))
    model.t_chunk_enc, model.t_chunk_dec = 16, 4
    assert model.build_chunk_start_end(129) == build_chunk_start_end(129, 16)
    assert model.build_chunk_start_end(33, decoder_mode=True) == build_chunk_start_end(33, 4)


def test_tiled_encode_equals_untiled_encode():
    """Invariant found by the survey: encoder tiling is exact, decoder tiling is not."""
    from oracle.vidtok_oracle import OracleModel, cfg_from_model_yaml
    d, meta = load_golden("tiny_kl_v11_tiled")
    sd, x = synth_weights(meta, d), synth_inputs(meta, d)
    om = OracleModel(cfg_from_model_yaml(meta["model"]), sd)
    _, _, h_full = om.encode(x, noise_fn=torch.zeros, return_pre=True)
    om.use_tiling, om.t_chunk_enc, om.t_chunk_dec, om.use_overlap = True, 16, 4, True
    _, _, h_tiled = om.encode(x, noise_fn=torch.zeros, return_pre=True)
    assert float((h_full - h_tiled).abs().max()) < 2e-5
