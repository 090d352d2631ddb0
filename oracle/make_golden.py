"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on seeded synthetic weights and inputs, and at the
same time pins oracle/vidtok_oracle.py against it (asserts agreement before writing anything).

Run in the authoring container only:   python oracle/make_golden.py [case ...]

Recipe (adapted so that nothing depends on module construction order):
  weights  = vidtok_b200.synth.synth_state_dict(reference state-dict shapes, seed=0)
  input    = vidtok_b200.synth.synth_clip(B, T, H, W, seed=1234)
  KL noise = torch.manual_seed(4321) immediately before encode(); the reference then calls
             torch.randn(mean.shape) once per regularizer invocation (distributions.py:17).
"""
from __future__ import annotations

import json
import os
import sys
import warnings

warnings.filterwarnings('ignore')

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.vidtok_oracle import OracleModel, cfg_from_model_yaml  # noqa: E402
from vidtok_b200.synth import synth_clip, synth_state_dict, weights_fingerprint  # noqa: E402

GOLDEN_DIR = os.environ.get("VIDTOK_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))   # override: regeneration test


def model_yaml(version="v1_0", reg="kl", ch=16, ch_mult=(1, 2, 4, 4), z=4, norm="layernorm", interp=None,
               levels=(8, 8, 8, 8, 8), causal=True, extra=None):
    mod = "model_3dcausal" + ("_v1_1" if version == "v1_1" else "")
    enc_cls, dec_cls = "EncoderCausal3DPadding", "DecoderCausal3DPadding"
    if not causal:   # configs/vidtok_kl_noncausal_488_4chn.yaml:12,30
        assert version == "v1_0"
        mod, enc_cls, dec_cls = "model_3dnoncausal", "Encoder3D", "Decoder3D"
    eng = "autoencoder" + ("_v1_1" if version == "v1_1" else "")
    ep = dict(double_z=(reg == "kl"), z_channels=z, in_channels=3, out_ch=3, ch=ch, ch_mult=list(ch_mult),
              time_downsample_factor=4, num_res_blocks=2, dropout=0.0, use_checkpoint=False,
              init_pad_mode="replicate", norm_type=norm, fix_encoder=False, fix_decoder=False)
    if interp is not None:
        ep["interpolation_mode"] = interp
    if extra:   # spatial_ds / spatial_us / tempo_ds / tempo_us / time_downsample_factor of the 444 / 288 / 888 configurations
        ep.update(extra)
    if reg == "kl":
        rc = {"target": "vidtok.modules.regularizers.DiagonalGaussianRegularizer"}
    else:
        rc = {"target": "vidtok.modules.regularizers.FSQRegularizer",
              "params": dict(levels=list(levels), entropy_loss_weight=0.1, entropy_loss_annealing_steps=2000,
                             entropy_loss_annealing_factor=3, commitment_loss_weight=0.25)}
    return {
        "target": f"vidtok.models.{eng}.AutoencodingEngine",
        "params": {
            "monitor": "val/rec_loss", "mode": "min", "ignore_keys": [],
            "encoder_config": {"target": f"vidtok.modules.{mod}.{enc_cls}", "params": ep},
            "decoder_config": {"target": f"vidtok.modules.{mod}.{dec_cls}",
                               "params": "${model.params.encoder_config.params}"},
            "regularizer_config": rc,
            "loss_config": {"target": "vidtok.modules.losses.GeneralLPIPSWithDiscriminator"},
        },
    }


# name -> (yaml kwargs, input shape (B,T,H,W), tiling (chunk or None), what to store)
CASES = {
    # non-causal family (model_3dnoncausal.py; 16-frame clips, configs/vidtok_kl_noncausal_488_4chn.yaml)
    "tiny_kl_nc": (dict(causal=False), (1, 16, 32, 32), None, "full"),
    "tiny_fsq_nc": (dict(causal=False, reg="fsq", z=5), (2, 16, 32, 32), None, "full"),
    "mid_kl_nc": (dict(causal=False, ch=64), (1, 16, 64, 64), None, "full"),
    # other compression ratios of the zoo: 4x4x4 (configs/vidtok_kl_causal_444_4chn.yaml:20-21), 2x8x8 (vidtok_kl_causal_288_8chn.yaml:20-22),
    # 8x8x8 v1.1 (vidtok_v1_1/vidtok_fsq_causal_888_32768_v1_1.yaml:21-23)
    "tiny_kl_444_v10": (dict(extra=dict(spatial_ds=[1, 2], spatial_us=[1, 2])), (1, 17, 32, 32), None, "full"),
    "tiny_kl_288_v10": (dict(z=8, extra=dict(tempo_ds=[1], tempo_us=[2], time_downsample_factor=2)), (1, 17, 32, 32), None, "full"),
    "tiny_fsq_888_v11": (dict(version="v1_1", reg="fsq", z=5, interp="trilinear",
                              extra=dict(tempo_ds=[0, 1, 2], tempo_us=[1, 2, 3], time_downsample_factor=8)), (1, 17, 32, 32), None, "full"),
    "tiny_kl_v10": (dict(), (1, 17, 32, 32), None, "full"),
    "tiny_kl_v10_t8": (dict(), (2, 8, 32, 32), None, "full"),
    "tiny_fsq_v10": (dict(reg="fsq", z=5), (2, 17, 32, 32), None, "full"),
    "tiny_kl_gn_v10": (dict(ch=128, norm="groupnorm"), (1, 17, 32, 32), None, "full"),
    "tiny_kl_41616_v10": (dict(ch_mult=(1, 2, 4, 4, 4)), (1, 17, 64, 64), None, "full"),
    "tiny_kl_v11": (dict(version="v1_1", z=16, interp="trilinear"), (1, 17, 32, 32), None, "full"),
    "tiny_kl_v11_tiled": (dict(version="v1_1", z=16, interp="trilinear"), (1, 49, 32, 32), 16, "full"),
    "tiny_fsq_v11_tiled": (dict(version="v1_1", reg="fsq", z=5, interp="trilinear"), (1, 33, 32, 32), 16, "full"),
    "mid_kl_v10": (dict(ch=64), (1, 17, 64, 64), None, "full"),
    "mid_fsq_v10": (dict(ch=64, reg="fsq", z=5), (1, 17, 64, 64), None, "full"),
    # BASELINE.json configs[0]: vidtok_kl_causal_488_4chn, 1x3x17x128x128
    "cfg1_kl_488_4chn": (dict(ch=128), (1, 17, 128, 128), None, "partial"),
    "cfg1_fsq_488_32768": (dict(ch=128, reg="fsq", z=5), (1, 17, 128, 128), None, "partial"),
}


def run_case(name: str):
    ykw, (B, T, H, W), chunk, store = CASES[name]
    my = model_yaml(**ykw)
    ref = ref_shim.build_reference_model(my)
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items() if k.startswith(("encoder.", "decoder."))}
    sd = synth_state_dict(shapes, seed=0)
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected and all(not m.startswith(("encoder.", "decoder.")) for m in missing), (missing, unexpected)
    x = synth_clip(B, T, H, W, seed=1234)

    if chunk is not None:
        ref.use_tiling = True
        ref.t_chunk_enc = chunk
        ref.t_chunk_dec = chunk // ref.encoder.time_downsample_factor
        ref.use_overlap = True

    # ---- the unmodified reference
    with torch.no_grad():
        torch.manual_seed(4321)
        z_ref, dec_ref, log_ref = ref(x)
        h_ref = None
        if chunk is None:
            if hasattr(ref, "_empty_causal_cached"):
                ref._empty_causal_cached(ref.encoder)
                ref._set_first_chunk(True)
            h_ref = ref.encoder(x)

    # ---- the oracle restatement
    cfg = cfg_from_model_yaml(my)
    om = OracleModel(cfg, sd)
    if chunk is not None:
        om.use_tiling, om.t_chunk_enc, om.t_chunk_dec, om.use_overlap = True, chunk, chunk // 4, True
    torch.manual_seed(4321)
    z_o, dec_o, log_o = om.forward(x)

    def maxabs(a, b):
        return float((a.double() - b.double()).abs().max())

    report = {"z": maxabs(z_ref, z_o), "dec": maxabs(dec_ref, dec_o)}
    assert dec_ref.shape == dec_o.shape, (dec_ref.shape, dec_o.shape)
    if T % 4 == 1 or ykw.get('version') == 'v1_1' or not ykw.get('causal', True):
        assert dec_ref.shape == x.shape, (dec_ref.shape, x.shape)
    assert report["z"] <= 2e-5 and report["dec"] <= 2e-5, report
    if "indices" in log_ref:
        assert torch.equal(log_ref["indices"], log_o["indices"]), "oracle FSQ indices differ from reference"
        assert log_ref["indices"].dtype == torch.int32
        # decode-from-indices path (README.md:344-348)
        with torch.no_grad():
            d2 = ref.decode(log_ref["indices"], decode_from_indices=True)
            if d2.shape[2] != x.shape[2]:
                d2 = d2[:, :, -x.shape[2]:]
        assert maxabs(d2, dec_ref) <= 1e-6
    else:
        report["kl"] = abs(float(log_ref["kl_loss"]) - float(log_o["kl_loss"])) / max(1.0, abs(float(log_ref["kl_loss"])))
        assert report["kl"] <= 1e-5, report

    out = {
        "z": z_ref.numpy(),
        "x_absum": np.float64(x.double().abs().sum()),
        "w_fingerprint": np.float64(weights_fingerprint(sd)),
    }
    if h_ref is not None:
        out["h"] = h_ref.numpy()  # encoder output before the regularizer
    if "indices" in log_ref:
        out["indices"] = log_ref["indices"].numpy()
    else:
        out["kl_loss"] = np.float32(log_ref["kl_loss"])
    if store == "full":
        out["x"] = x.numpy()
        out["dec"] = dec_ref.numpy()
    else:
        frames = [0, T // 2, T - 1]
        out["dec_frames"] = np.array(frames)
        out["dec_sel"] = dec_ref[:, :, frames].numpy()
        out["dec_frame_mean"] = dec_ref.double().mean(dim=(0, 1, 3, 4)).numpy()
        out["dec_frame_absmean"] = dec_ref.double().abs().mean(dim=(0, 1, 3, 4)).numpy()
    meta = {
        "case": name, "model": my, "input": [B, 3, T, H, W], "tiling_chunk": chunk,
        "weights_seed": 0, "input_seed": 1234, "noise_seed": 4321,
        "shapes": {k: list(v) for k, v in shapes.items()},
        "oracle_vs_reference_maxabs": report,
        "torch": torch.__version__,
    }
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(f"[golden] {name}: oracle-vs-reference {report}  -> {name}.npz", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    names = sys.argv[1:] or list(CASES.keys())
    for n in names:
        run_case(n)
