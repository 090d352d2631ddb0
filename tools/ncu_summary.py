"""Summarise ncu outputs from gpurun_out/ into profiles/ (tracked).
  python tools/ncu_summary.py <round-tag>     e.g. r1
Reads gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum launch list) and
gpurun_out/prof_conv_tc_<tag>.ncu-rep (ncu --set full capture of the dominant kernel)."""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list
path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(path):
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[iu]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        name = r[ik].split("(")[0].replace("void ", "").replace("vt::<unnamed>::", "").replace("vt::(anonymous namespace)::", "")
        agg[name][0] += 1
        agg[name][1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"launches_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list, round {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n"
                "(model build + weight packing + PSNR check + device-resident and end-to-end steps of 8 clips 17x256x256, bf16).  Times are cold-cache and serialised:\n"
                "compare SHARES, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |\n")
        f.write(f"\ntotal {tot / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches\n")
    print("wrote", f"profiles/launches_{tag}.md")

# ---- full capture of conv_tc
rep = os.path.join(ROOT, "gpurun_out", f"prof_conv_tc_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(out_dir, f"ncu_conv_tc_{tag}.md"), "w") as f:
        f.write(f"# ncu --set full, conv_tc_kernel, round {tag}\n\n`ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 100 -c 4`\n"
                "over `python tools/ncu_target.py 8 1` (launches 101-104 of the kernel: decoder level-0/1 layers).\n\n")
        for r in rows[2:]:
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for w in want:
                if w in idx:
                    f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
            f.write("\n")
    print("wrote", f"profiles/ncu_conv_tc_{tag}.md")
    # machine-readable digest for bench.py's roofline.traffic
    import json
    def fval(r, name):
        try:
            v = float(r[idx[name]].replace(",", ""))
        except Exception:
            return None
        u = units[idx[name]]
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    launches = []
    for r in rows[2:]:
        rd, wr, ms = fval(r, "dram__bytes_read.sum"), fval(r, "dram__bytes_write.sum"), fval(r, "gpu__time_duration.sum")
        tp = fval(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
        launches.append({"ms": ms, "dram_read_bytes": rd, "dram_write_bytes": wr, "tensor_pipe_active_pct": tp})
    json.dump({"kernel": "conv_tc_kernel", "capture": f"ncu --set full -k regex:conv_tc_kernel over tools/ncu_target.py 8 1 (round {tag})", "launches": launches},
              open(os.path.join(out_dir, f"ncu_conv_tc_{tag}.json"), "w"), indent=1)
