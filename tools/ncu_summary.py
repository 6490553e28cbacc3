#!/usr/bin/env python
"""Summarise ncu outputs for profiles/: a launch list (`--metrics gpu__time_duration.sum --csv --log-file`) and the raw
page of one `--set full` capture (`ncu -i X.ncu-rep --page raw --csv`).  Usage:
    python tools/ncu_summary.py launches.csv raw.csv bench.json > profiles/rNN_ncu.md
    python tools/ncu_summary.py --traffic-json raw.csv profiles/ncu_traffic.json c4_jacket10k profiles/rNN_raw.csv.gz"""
import csv
import json
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("jk::", "")
    return name[:90]


def launch_table(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v *= 1e-3
        elif r[ui] in ("ms", "msecond"):
            v *= 1e3
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {us / 1e3:.3f} | {us / n:.2f} | {us / tot:.3f} |")
    return "\n".join(out)


METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem / block", "launch__shared_mem_per_block_dynamic"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 sectors % of peak", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe active %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("DMMA pipe active %", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("LSU data pipe % (elapsed)", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall math throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall sleeping", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"),
]


def full_table(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(short(r[ki]), r) for r in rows[2:]]
    out = ["| metric | " + " | ".join(f"`{c[0][:28]}`" for c in cols) + " | unit |", "|---|" + "---:|" * len(cols) + "---|"]
    for label, m in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        vals = []
        for _, r in cols:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        out.append(f"| {label} | " + " | ".join(vals) + f" | {units[i]} |")
    return "\n".join(out)


def traffic_json(raw_path, out_path, workload, source):
    """profiles/ncu_traffic.json: dram read + write bytes per launch of every captured kernel (mean over its launches) --
    bench.py fills roofline.traffic from this file, never from a constant."""
    import os
    rows = list(csv.reader(open(raw_path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = {}
    for r in rows[2:]:
        name = re.sub(r"<.*", "", short(r[ki]))
        try:
            b = float(r[ri].replace(",", "")) * scale.get(units[ri], 1.0) + float(r[wi].replace(",", "")) * scale.get(units[wi], 1.0)
        except ValueError:
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += b
    data = json.load(open(out_path)) if os.path.isfile(out_path) else {}
    # the capture covers exactly one step: bytes_per_step = all launches of the kernel in that step
    data[workload] = {k: {"bytes_per_launch": v[1] / v[0], "launches_in_step": v[0], "bytes_per_step": v[1], "source": source} for k, v in agg.items()}
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)


def main():
    if sys.argv[1] == "--traffic-json":       # tools/ncu_summary.py --traffic-json raw.csv profiles/ncu_traffic.json workload source
        traffic_json(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
        return
    launches, raw = sys.argv[1], sys.argv[2]
    print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
    print(launch_table(launches))
    print("\n## `ncu --set full --clock-control none --import-source on` (one launch per column, in launch order)\n")
    print(full_table(raw))
    if len(sys.argv) > 3:
        d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
        print("\n## Bench line of the same build (no ncu)\n")
        print(f"{d['value']:.0f} {d['unit']}, {d['ms_per_step']:.3f} ms/step, {d['gpu_launches']} launches in the timed region; stage timers (ms): "
              + json.dumps({k: round(v, 3) for k, v in d["stage_ms"].items()}))
        print("\nroofline: " + json.dumps(d["roofline"]))


if __name__ == "__main__":
    main()
