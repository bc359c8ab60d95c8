#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as fp:
        lines = [ln for ln in fp if not ln.startswith("==")]
    reader = csv.DictReader(lines)
    for r in reader:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0,
                 "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], val * scale))
    agg = OrderedDict()
    for name, ms in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        c = agg.setdefault(short, [0, 0.0])
        c[0] += 1
        c[1] += ms
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total:.3f} ms of device time (cold-cache, serialised under ncu)")
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {n:8d} {ms:10.3f} {ms / n:9.4f} {100 * ms / total:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
