#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of time per kernel.

usage: python profiles/summarize_launches.py launches.csv [launches_per_step_hint] > summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        name = r["Kernel Name"]
        rows.append((name, us, r["Grid Size"], r["Block Size"]))
    agg = defaultdict(list)
    shape = {}
    for name, us, g, b in rows:
        key = name.split("(")[0][:70]
        agg[key].append(us)
        shape[key] = (g, b)
    total = sum(us for _, us, _, _ in rows)
    print(f"launches {len(rows)}  total {total:.1f} us")
    for key, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        v2 = sorted(v)
        print(f"{100 * sum(v) / total:5.1f}%  n={len(v):4d} avg={sum(v) / len(v):8.2f}us med={v2[len(v2) // 2]:8.2f} "
              f"max={v2[-1]:8.2f}  grid={shape[key][0]} block={shape[key][1]}  {key}")


if __name__ == "__main__":
    main(sys.argv[1])
