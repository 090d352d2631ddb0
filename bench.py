#!/usr/bin/env python
"""Benchmark of the tokenizer hot path: frames/sec for full encode -> regularize -> decode of
vidtok_kl_causal_488_4chn on synthetic 17x256x256 clips (BASELINE.json metric, configs[1]).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch of 8 clips per GPU (weak scaling).  `value` is timed with CUDA
events with the inputs already resident in HBM; `e2e` goes through the public Python API
(vidtok.models.autoencoder.AutoencodingEngine.forward) from pinned host memory and back.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "frames/sec encode+decode, kl_causal_488 17x256x256"
FLOPS_PER_CLIP = 20.691e12  # BASELINE.md section 2 (FlopCounterMode over the reference graph)
T_CLIP, H_CLIP, W_CLIP = 17, 256, 256


def model_cfg(ch=128):
    """configs/vidtok_kl_causal_488_4chn.yaml:1-36 (model section)."""
    ep = dict(double_z=True, z_channels=4, in_channels=3, out_ch=3, ch=ch, ch_mult=[1, 2, 4, 4],
              time_downsample_factor=4, num_res_blocks=2, dropout=0.0, use_checkpoint=False,
              init_pad_mode="replicate", norm_type="layernorm", fix_encoder=False, fix_decoder=False)
    return {
        "target": "vidtok.models.autoencoder.AutoencodingEngine",
        "params": {
            "monitor": "val/rec_loss", "mode": "min", "ignore_keys": [],
            "encoder_config": {"target": "vidtok.modules.model_3dcausal.EncoderCausal3DPadding", "params": ep},
            "decoder_config": {"target": "vidtok.modules.model_3dcausal.DecoderCausal3DPadding", "params": dict(ep)},
            "regularizer_config": {"target": "vidtok.modules.regularizers.DiagonalGaussianRegularizer"},
            "loss_config": {"target": "vidtok.modules.losses.GeneralLPIPSWithDiscriminator"},
        },
    }


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1431.0))), "hbm_gbs": float(d.get("hbm_gbs", 6568.0)),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"}


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU reference (oracle port of the reference's PyTorch CPU path)
# --------------------------------------------------------------------------------------------------
def oracle_model(sd):
    from oracle.vidtok_oracle import OracleModel, cfg_from_model_yaml
    return OracleModel(cfg_from_model_yaml(model_cfg()), sd)


def cpu_forward_timed(om, x, noise_seed=4321):
    torch.manual_seed(noise_seed)
    t0 = time.perf_counter()
    z, dec, _ = om.forward(x)
    return time.perf_counter() - t0, dec


def pick_cpu_threads(om):
    """The reference would run with torch's default (all host cores).  On cgroup-limited hosts that oversubscribes
    badly (128 visible cores, far fewer usable), so probe a few thread counts on a tiny clip and keep the fastest."""
    from vidtok_b200.synth import synth_clip
    cores = os.cpu_count() or 1
    try:
        cores = min(cores, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    x = synth_clip(1, T_CLIP, 32, 32)
    best_t, best_n = None, cores
    for n in sorted({cores, max(cores // 2, 1), max(cores // 4, 1), 32, 16, 8}):
        if n > cores:
            continue
        torch.set_num_threads(n)
        t, _ = cpu_forward_timed(om, x)
        t2, _ = cpu_forward_timed(om, x)
        t = min(t, t2)
        if best_t is None or t < best_t:
            best_t, best_n = t, n
        elif t > 1.5 * best_t:  # more threads only oversubscribe from here on
            break
    torch.set_num_threads(best_n)
    return best_n


def pick_cpu_sample(om, budget_s: float, steps: int):
    """Largest sample clip (17 x S x S, S in 256/128/64) whose `steps` forwards fit the budget, from a 64x64 probe."""
    from vidtok_b200.synth import synth_clip
    t_probe, _ = cpu_forward_timed(om, synth_clip(1, T_CLIP, 64, 64))
    t_probe2, _ = cpu_forward_timed(om, synth_clip(1, T_CLIP, 64, 64))
    t64 = min(t_probe, t_probe2)
    for S in (256, 128, 64):
        if t64 * (S / 64) ** 2 * steps <= budget_s:
            return S
    return 64


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vidtok_b200.compat_util import instantiate_from_config  # noqa: F401  (manifest source for weight shapes)
    from vidtok_b200.engine import NativeModel, TokenizerSpec
    from vidtok_b200.synth import synth_clip, synth_state_dict
    spec = TokenizerSpec.from_params(model_cfg()["params"]["encoder_config"]["params"], 0)
    sd = synth_state_dict(dict(NativeModel(spec).manifest()), seed=0)
    om = oracle_model(sd)
    cores = pick_cpu_threads(om)
    total = args.steps + args.warmup
    S = pick_cpu_sample(om, budget_s=240.0, steps=total)
    x = synth_clip(1, T_CLIP, S, S)
    for _ in range(args.warmup):
        cpu_forward_timed(om, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_forward_timed(om, x)
    el = time.perf_counter() - t0
    scale = (S * S) / float(H_CLIP * W_CLIP)
    fps = T_CLIP * args.steps / el * scale
    sample = f"1 clip 3x{T_CLIP}x{S}x{S} per step on {cores} host threads of {os.cpu_count()} visible (oracle port of the reference PyTorch CPU path)"
    if S != 256:
        sample += f"; value scaled by the pixel ratio {scale:.4f} to 256x256-frame units"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "vidtok_kl_causal_488_4chn: clips 17x256x256 (reference CPU path, fp32)", "sample": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    from vidtok_b200 import _native as N
    from vidtok_b200 import dist as vdist
    from vidtok_b200.compat_util import instantiate_from_config
    from vidtok_b200.synth import synth_clip, synth_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference for the CPU path)")
    rank, world, local = vdist.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = args.batch
    model = instantiate_from_config(model_cfg())
    sd = synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    model.precision = "bf16"
    lib = N.lib()

    x_host = synth_clip(B, T_CLIP, H_CLIP, W_CLIP, seed=1234 + rank).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x_dev = x_host.to(dev)
    torch.manual_seed(4321)

    def step_resident():
        with torch.no_grad():
            return model(x_dev)

    # End-to-end leg: the call a user makes (model(x) on the current stream) with every step's input coming from pinned host
    # memory and every step's reconstruction going back to pinned host memory.  The copies run on a second stream, double
    # buffered: H2D of step k+1 and D2H of step k-1 overlap the kernels of step k (each step's copies stay inside the timed
    # region: the first H2D and the last D2H are not hidden).
    copy_stream = torch.cuda.Stream(device=dev)
    in_bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]

    def run_e2e(steps):
        main = torch.cuda.current_stream(dev)
        ready = [torch.cuda.Event(), torch.cuda.Event()]   # input buffer i holds its step's clip
        freed = [torch.cuda.Event(), torch.cuda.Event()]   # the step that read input buffer i has finished
        dec = None
        with torch.no_grad():
            with torch.cuda.stream(copy_stream):
                in_bufs[0].copy_(x_host, non_blocking=True)
                ready[0].record(copy_stream)
            for k in range(steps):
                cur = k & 1
                main.wait_event(ready[cur])
                _, dec, _ = model(in_bufs[cur])
                freed[cur].record(main)
                dec.record_stream(copy_stream)
                with torch.cuda.stream(copy_stream):
                    if k + 1 < steps:
                        if k >= 1:
                            copy_stream.wait_event(freed[cur ^ 1])
                        in_bufs[cur ^ 1].copy_(x_host, non_blocking=True)
                        ready[cur ^ 1].record(copy_stream)
                    copy_stream.wait_event(freed[cur])
                    out_host.copy_(dec, non_blocking=True)
        main.wait_stream(copy_stream)
        return dec

    # the clock sampler starts BEFORE the warm-up: nvidia-smi takes ~1 s to initialise NVML, and doing that inside the
    # timed region cost the first steps ~10 % (profiles/notes_r1.md); its samples cover warm-up + timed steps, all under load
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize(dev)

    # ---- device-timed region (inputs resident in HBM)
    vdist.barrier()
    torch.cuda.synchronize(dev)
    lib.vt_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        z, dec, log = step_resident()
    e1.record()
    torch.cuda.synchronize(dev)
    vdist.barrier()
    launches = int(lib.vt_launch_count(0))
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    ms = float(vdist.allreduce_max(ms)[0])
    frames = world * B * T_CLIP * args.steps
    value = frames / (ms / 1e3)

    # ---- end to end through the public API with host buffers
    run_e2e(2)
    torch.cuda.synchronize(dev)
    vdist.barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    dec = run_e2e(args.steps)
    e3.record()
    torch.cuda.synchronize(dev)
    vdist.barrier()
    ms_e2e = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
    ms_e2e = float(vdist.allreduce_max(ms_e2e)[0])
    e2e_value = frames / (ms_e2e / 1e3)

    # ---- the one collective: global PSNR(input, reconstruction) from per-rank partial sums (NCCL all-reduce)
    psnr_b200 = vdist.global_psnr(vdist.psnr_partial(x_dev, dec))

    # ---- per-kernel attribution of one step (CUDA events around every launch, on the launch stream)
    peaks = load_peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_conv_tc_r1.json")
    if os.path.exists(tpath):  # dram__bytes_read+write per launch from the committed `ncu --set full` capture
        try:
            cap = json.load(open(tpath))
            vals = [l["dram_read_bytes"] + l["dram_write_bytes"] for l in cap["launches"] if l.get("dram_read_bytes") is not None]
            traffic = {"bytes_per_launch_avg": sum(vals) / len(vals), "launches_captured": len(vals), "source": "profiles/ncu_conv_tc_r1.json"}
        except Exception:
            traffic = None
    roof = None
    prof = {}
    if rank == 0:
        lib.vt_profile_start()
        step_resident()
        buf = __import__("ctypes").create_string_buffer(1 << 16)
        n = lib.vt_profile_stop(buf, len(buf))
        prof = json.loads(buf.value.decode()) if n > 0 else {}
        tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0] if prof else None
        if dom is not None:
            d = prof[dom]
            ach = d["flops"] / (d["ms"] / 1e3) / 1e12 if d["flops"] > 0 else d["bytes"] / (d["ms"] / 1e3) / 1e9
            bound = "tensor" if d["flops"] > 0 and dom.startswith("conv") else "hbm"
            peak = peaks["tflops"] if bound == "tensor" else peaks["hbm_gbs"]
            roof = {"kernel": dom, "bound": bound, "achieved": ach, "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                    "frac": ach / peak, "traffic": traffic, "launches_per_step": d["launches"],
                    "avg_launch_ms": d["ms"] / max(d["launches"], 1), "share_of_step": d["ms"] / tot_ms,
                    "algorithmic_flops_per_step": d["flops"], "peak_source": peaks["source"],
                    "whole_path": {"achieved": FLOPS_PER_CLIP * world * B * args.steps / (ms / 1e3) / 1e12 / world, "unit": "TFLOP/s per GPU",
                                   "frac": FLOPS_PER_CLIP * B * args.steps / (ms / 1e3) / 1e12 / peaks["tflops"]},
                    "kernels_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}

    # ---- CPU baseline + PSNR parity on clip 0 (rank 0, N == 1 only)
    cpu = None
    psnr = {"b200_all_clips": psnr_b200}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        om = oracle_model(sd)
        cores = pick_cpu_threads(om)
        S = pick_cpu_sample(om, budget_s=40.0, steps=1)
        xs = x_host[:1] if S == 256 else synth_clip(1, T_CLIP, S, S)
        el, dec_ref = cpu_forward_timed(om, xs)
        scale = (S * S) / float(H_CLIP * W_CLIP)
        sample = f"1 clip 3x{T_CLIP}x{S}x{S}, 1 forward, fp32, {cores} host threads of {os.cpu_count()} visible (oracle port of the reference PyTorch CPU path)"
        if S != 256:
            sample += f"; value scaled by the pixel ratio {scale:.4f} to 256x256-frame units"
        cpu = {"value": T_CLIP / el * scale, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
        # PSNR gate on the same clip, same weights, same noise
        with torch.no_grad():
            torch.manual_seed(4321)
            _, dec_g, _ = model(xs.to(dev))
        from vidtok_b200.dist import psnr_partial
        pg = psnr_partial(xs, dec_g.cpu())
        pr = psnr_partial(xs, dec_ref)
        psnr.update({"clip": f"3x{T_CLIP}x{S}x{S}", "b200_bf16": float(pg[0] / pg[1]), "reference_cpu_fp32": float(pr[0] / pr[1]),
                     "abs_diff_db": abs(float(pg[0] / pg[1]) - float(pr[0] / pr[1])), "gate_db": 0.01})

    if rank == 0:
        nbytes = x_host.numel() * x_host.element_size()
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "vidtok_kl_causal_488_4chn bf16: batch 8 clips 17x256x256 per GPU (BASELINE.json configs[1])",
                       "clips_per_gpu": B, "parallelism": f"dp{world} (clips sharded, no data-path collective)",
                       "weights": "random (synth_state_dict seed 0)", "l2": "per-step activations are GBs, far larger than the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": nbytes + 4 * B * 4 * 5 * 32 * 32, "d2h_bytes_per_step": nbytes,
                    "pipeline": "every step copies its clips from pinned host memory and its reconstruction back; the copies run on a "
                                "second stream, double buffered against the previous / next step's kernels"},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
            "psnr": psnr,
        }
        print(json.dumps(line), flush=True)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
