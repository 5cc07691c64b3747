#!/usr/bin/env python
"""Summarise ncu outputs for profiles/: (1) a launch list CSV -> time share per kernel,
(2) a .ncu-rep -> the roofline-relevant raw metrics.  Usage:
    python tools/ncu_summary.py launches gpurun_out/launches_X.csv
    python tools/ncu_summary.py rep gpurun_out/prof_X.ncu-rep
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    mu = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        t = float(r[mv].replace(",", ""))
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)   # -> us
        a = agg[short(r[kn])]
        a[0] += 1
        a[1] += t
        total += t
    print(f"launches: {sum(a[0] for a in agg.values())}, summed device time {total / 1e3:.3f} ms (ncu: cold-cache, serialised; compare SHARES)")
    print(f"{'kernel':90s} {'n':>6s} {'ms':>9s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:90]:90s} {n:6d} {t / 1e3:9.3f} {100 * t / total:6.1f}%")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:150])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:62s} {r[i]:>16s} {units[i]}")
        try:
            tr = float(r[hdr.index("dram__bytes_read.sum")]) + float(r[hdr.index("dram__bytes_write.sum")])
            u = units[hdr.index("dram__bytes_read.sum")]
            print(f"  {'traffic = dram read + write':62s} {tr:16.3f} {u}")
        except Exception:
            pass


def traffic(path, key, algorithmic_bytes, match=""):
    """traffic <rep> <op:fmt:in->out> <algorithmic bytes of the captured launch> [kernel-name substring]: writes the DRAM
    bytes of the first matching launch into profiles/r2_traffic.json (what bench.py's roofline.traffic is read from)."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if match and match not in name:
            continue
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        rd, wr = float(r[ir]) * scale[units[ir]], float(r[iw]) * scale[units[iw]]
        dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json")
        table = json.load(open(dst)) if os.path.exists(dst) else {}
        table[key] = {"kernel": short(name), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                      "algorithmic_bytes": float(algorithmic_bytes), "source": f"ncu --set full --clock-control none, {os.path.basename(path)}",
                      "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")])}
        json.dump(table, open(dst, "w"), indent=1, sort_keys=True)
        print(json.dumps(table[key]))
        return
    raise SystemExit("no matching launch")


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2])
