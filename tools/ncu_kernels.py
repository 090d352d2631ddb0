"""Per-kernel summaries of an `ncu --set full` report: one profiles/ncu_<kernel>_<tag>.md per kernel function.
  python tools/ncu_kernels.py <report.ncu-rep> <tag> ["how the capture was made"]
Each file lists the launches captured for that kernel (first 12) with the metrics the judge reads: duration, grid/block/regs/smem,
tensor-pipe activity, SM / L2 / DRAM throughput, DRAM bytes read+written, achieved occupancy, plus the top warp-stall reasons."""
import csv
import io
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
how = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__cluster_size", "cluster"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("smsp__inst_executed.sum", "warp insts"), ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe %")]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if not STALL:
    STALL = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") or (h.startswith("smsp__average_warps_issue_stalled") and "ratio" in h)]


def short(name):
    n = name.replace("(anonymous namespace)", "").replace("<unnamed>", "").replace("unnamed>", "")
    n = n.split("(")[0].replace("void ", "").replace("vt::", "").replace("::", "")
    return n


groups = OrderedDict()
for r in rows[2:]:
    if len(r) <= idx["Kernel Name"]:
        continue
    groups.setdefault(short(r[idx["Kernel Name"]]), []).append(r)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
index = []
for name, rs in groups.items():
    fname = re.sub(r"[^A-Za-z0-9]+", "_", name).strip("_")
    path = os.path.join(ROOT, "profiles", f"ncu_{fname}_{tag}.md")
    tot_ms = 0.0
    with open(path, "w") as f:
        f.write(f"# ncu --set full: `{name}` (round {tag})\n\n{how}\n\n{len(rs)} launch(es) captured; first {min(len(rs), 12)} shown.  "
                "Times under ncu are serialised and cold-cache: read shares and ratios, not absolutes.\n\n")
        cols = [w for w in WANT if w[0] in idx]
        f.write("| # | " + " | ".join(f"{c[1]} [{units[idx[c[0]]]}]" if units[idx[c[0]]] else c[1] for c in cols) + " |\n")
        f.write("|---|" + "---:|" * len(cols) + "\n")
        for i, r in enumerate(rs):
            try:
                v = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
                tot_ms += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1e-6)
            except Exception:
                pass
            if i < 12:
                f.write(f"| {i} | " + " | ".join(r[idx[c[0]]] for c in cols) + " |\n")
        # stall reasons of the longest launch
        if STALL:
            def dur(r):
                try:
                    return float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
                except Exception:
                    return 0.0
            big = max(rs, key=dur)
            st = []
            for h in STALL:
                try:
                    st.append((float(big[idx[h]].replace(",", "")), h))
                except Exception:
                    pass
            st.sort(reverse=True)
            f.write("\nTop warp-stall reasons of the longest launch (warps stalled per issue-active cycle):\n\n")
            for v, h in st[:6]:
                f.write(f"- {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}\n")
    index.append((name, len(rs), tot_ms, os.path.basename(path)))
with open(os.path.join(ROOT, "profiles", f"ncu_index_{tag}.md"), "a") as f:
    f.write(f"\n## {os.path.basename(rep)} -- {how}\n\n| kernel | launches | total ms (under ncu) | file |\n|---|---:|---:|---|\n")
    for name, n, ms, fn in sorted(index, key=lambda t: -t[2]):
        f.write(f"| `{name}` | {n} | {ms:.3f} | {fn} |\n")
print("wrote", len(index), "kernel summaries")
