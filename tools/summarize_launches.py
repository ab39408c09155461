#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, frames):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name).replace("unnamed>::", "").replace("rmem::", "").replace("<", "<")
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] == "us":
            ns *= 1e3
        rows.append((name, r["Grid Size"], ns))
    agg = defaultdict(lambda: [0, 0.0])
    for name, grid, ns in rows:
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"launches: {len(rows)}  frames: {frames}  launches/frame: {len(rows)/frames:.1f}  "
          f"sum of kernel time: {total/1e3/frames:.1f} us/frame (serialised, cold-cache)")
    print(f"{'kernel':60s} {'n/frame':>8s} {'us/frame':>10s} {'share':>7s} {'avg us':>8s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n/frames:8.1f} {ns/1e3/frames:10.1f} {100*ns/total:6.1f}% {ns/1e3/n:8.2f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
