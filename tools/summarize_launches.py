#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, summed device time and share per kernel.

    python tools/summarize_launches.py gpurun_out/launches_r1d.csv > profiles/r1_launches_table.md
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((re.sub(r"\(.*", "", r[ki]), ms))
    tot = sum(ms for _, ms in rows)
    by = defaultdict(lambda: [0, 0.0])
    for k, ms in rows:
        by[k][0] += 1
        by[k][1] += ms
    print(f"launches captured: {len(rows)}, summed kernel time {tot:.1f} ms\n")
    print("| kernel | launches | summed ms | share |\n|---|---:|---:|---:|")
    for k, (n, ms) in sorted(by.items(), key=lambda kv: -kv[1][1])[:18]:
        print(f"| `{k[:80]}` | {n} | {ms:.2f} | {100 * ms / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
