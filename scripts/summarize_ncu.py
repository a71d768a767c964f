#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (tracked; gpurun_out/ is scratch).

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
    python scripts/summarize_ncu.py kernels  gpurun_out/prof.ncu-rep > profiles/rNN_kernels.md
    python scripts/summarize_ncu.py metrics  gpurun_out/ncu.csv      > profiles/rNN_ncu.md
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe inst % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: fixed-latency wait"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local (spill) load sectors"),
]


def short(name):
    name = name.replace("void ", "").replace("tplb::", "")
    name = name.split("(")[0]
    return name.replace("<unnamed>::Model", "Model")


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        agg.setdefault(short(r[ki]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.1f} | {sum(v) / tot:.1%} |")
    print(f"\ntotal {tot / 1e3:.3f} ms over {sum(len(v) for v in agg.values())} launches "
          "(ncu serialises launches and runs them cold-cache: compare shares, not absolutes)")


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"### `{short(r[ki])}`\n\n| metric | value |\n|---|---:|")
        for key, label in KEEP:
            if key in hdr:
                i = hdr.index(key)
                val = r[i]
                try:
                    val = f"{float(val.replace(',', '')):,.3f}".rstrip("0").rstrip(".")
                except ValueError:
                    pass
                print(f"| {label} (`{key}`) | {val} {units[i]} |")
        print()


def metrics(path):
    """One table for a `ncu --metrics ... --csv --page raw --log-file` capture (one row per launch)."""
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[h], rows[h + 1]
    ki = hdr.index("Kernel Name")
    cols = [(k, lab) for k, lab in KEEP if k in hdr]
    print("| # | kernel | " + " | ".join(lab for _, lab in cols) + " |")
    print("|---:|---|" + "---:|" * len(cols))
    for n, r in enumerate(rows[h + 2:]):
        if len(r) < len(hdr):
            continue
        vals = []
        for k, _ in cols:
            i = hdr.index(k)
            try:
                v = float(r[i].replace(",", ""))
                vals.append((f"{v:,.0f}" if abs(v) >= 1000 else f"{v:.2f}") + (f" {units[i]}" if units[i] not in ("", "%", "inst", "register/thread") else ""))
            except ValueError:
                vals.append(r[i])
        print(f"| {n} | `{short(r[ki]).split('<')[0]}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels, "metrics": metrics}[sys.argv[1]](sys.argv[2])
