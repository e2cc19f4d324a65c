#!/usr/bin/env python
"""Summarise ncu output for profiles/: (a) a launch list CSV (--metrics gpu__time_duration.sum) -> per-kernel share table,
(b) a --set full .ncu-rep -> table of the metrics the design discussion uses.

    python tools/ncu_summary.py launches gpurun_out/launches_x.csv "title"      >> profiles/rNN_launch_summary.md
    python tools/ncu_summary.py full gpurun_out/prof_x.ncu-rep "title"          >  profiles/rNN_ncu_x.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("fv::", "")[:80]


def launches(path: str, title: str) -> None:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for r in rows:
        if r is hdr or len(r) <= iv or r[iv] in ("", "Metric Value"):
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu].strip(), 1e-3)
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += v * scale
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"## {title}\n")
    print(f"`{path}`: {n} launches, {tot:.1f} us total (ncu `gpu__time_duration.sum`, cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    print()


def full(path: str, title: str) -> None:
    if path.endswith(".csv"):   # raw page already exported on the GPU box (tools/run_profiles.sh)
        out = open(path, errors="replace").read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n")
    cols = rows[2:]
    names = [short(r[hdr.index("Kernel Name")]) for r in cols]
    print("| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(cols))) + " | unit |")
    print("|---|" + "---:|" * len(cols) + "---|")
    print("| kernel | " + " | ".join(f"`{n}`" for n in names) + " | |")
    for m in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        vals = []
        for r in cols:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| `{m}` | " + " | ".join(vals) + f" | {units[i]} |")
    print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[2])
