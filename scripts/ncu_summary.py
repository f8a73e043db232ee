#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python scripts/ncu_summary.py <tag>      # e.g. r1a

  gpurun_out/launches.csv      -> profiles/<tag>_launches.md   (per-kernel count / total / share)
  gpurun_out/prof_*.ncu-rep    -> profiles/<tag>_<name>.md     (key metrics per captured launch)
"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_membar.ratio",
        "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
        "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_selected.ratio"]


def launches(tag):
    p = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.exists(p):
        return
    lines = [ln for ln in open(p) if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, --clock-control none\n\n")
        f.write("Cold-cache, serialised launches: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | avg ms | share | grid | block |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / a[0]:.4f} | {a[1] / tot:.4f} | {a[2]} | {a[3]} |\n")
    print("wrote", f"{tag}_launches.md")


def reports(tag):
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(os.path.join(OUT, f"{tag}_{name}.md"), "w") as f:
            f.write(f"# ncu --set full capture ({tag}, {name}); numbers under a profiler are not bench values\n\n")
            for r in rows[2:]:
                d = dict(zip(hdr, r))
                f.write(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n")
                f.write("| metric | value | unit |\n|---|---|---|\n")
                for k in KEYS:
                    if k in d:
                        f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
                try:
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                    tr = sum(float(d[k].replace(",", "")) * scale[units[hdr.index(k)]]
                             for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                    f.write(f"| dram traffic (read+write) | {tr / 1e9:.3f} | Gbyte |\n")
                except Exception:
                    pass
                f.write("\n")
        print("wrote", f"{tag}_{name}.md")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "rX"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    reports(tag)
