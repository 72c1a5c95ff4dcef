#!/usr/bin/env python
"""Summarises ncu captures brought back in gpurun_out/ into profiles/ (tracked).

  python tools/summarize_ncu.py r01

* launches_<tag>.csv  -> profiles/launches_<tag>.md : per-kernel launch count, total and share of
  device time for the bench command (cold-cache, serialised: compare SHARES, not absolutes).
* prof_<kernel>_<tag>.ncu-rep -> profiles/<kernel>_<tag>.txt : the metrics the roofline uses
  (duration, dram bytes, achieved occupancy, registers, pipe utilisation, top stall reasons).
"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:90]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        val_us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit.startswith("us") else val * 1e3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val_us
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"launches_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list `{tag}`: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: "
                "read the SHARES).\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {cnt} | {us:.1f} | {100 * us / total:.2f}% |\n")
        f.write(f"\ntotal {total / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches\n")
    print("wrote launches", len(agg), "kernels")


def full(tag):
    for fn in sorted(os.listdir(OUT)):
        m = re.match(rf"prof_(.+)_{tag}\.(ncu-rep|raw\.csv)$", fn)
        if not m:
            continue
        kern = m.group(1)
        if m.group(2) == "raw.csv":   # already exported on the GPU box (the .ncu-rep files exceed the copy-back limit)
            text = open(os.path.join(OUT, fn), errors="replace").read()
            text = text[text.find('"ID"'):]
            class R: stdout = text; stderr = ""
            r = R()
        else:
            r = subprocess.run(["ncu", "-i", os.path.join(OUT, fn), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            print("no data in", fn, r.stderr[:200])
            continue
        header, units, vals = rows[0], rows[1], rows[2]
        d = dict(zip(header, zip(units, vals)))
        with open(os.path.join(PROF, f"{kern}_{tag}.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none, kernel {d.get('Kernel Name', ('', '?'))[1]}\n")
            f.write(f"# grid {d.get('Grid Size', ('', '?'))[1]} block {d.get('Block Size', ('', '?'))[1]}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:80s} {d[k][1]:>18s} {d[k][0]}\n")
            f.write("\n# warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active / pcsamp)\n")
            stalls = [(k, d[k][1]) for k in d if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
            for k, v in sorted(stalls, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:10]:
                f.write(f"{k:100s} {v}\n")
        print("wrote", kern)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    full(tag)
