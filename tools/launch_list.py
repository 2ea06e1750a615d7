#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list into the text kept under
profiles/: per-kernel aggregate (time, share, launches) followed by every launch in order.

    python tools/launch_list.py gpurun_out/launches.csv "header line(s)" > profiles/rNN_launches_bench.txt"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    header = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    cols = next(rd)
    ix = {c: i for i, c in enumerate(cols)}
    for r in rd:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((int(r[ix["ID"]]), ms, r[ix["Grid Size"]], r[ix["Block Size"]], r[ix["Kernel Name"]]))
    total = sum(r[1] for r in rows)
    agg = defaultdict(lambda: [0.0, 0])
    for _, ms, _, _, name in rows:
        agg[name][0] += ms
        agg[name][1] += 1
    if header:
        print(header)
    print(f"# {len(rows)} launches, {total:.3f} ms of kernel time in total (cold-cache, serialised: compare shares, not absolutes)\n")
    print("## aggregate")
    for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:10.3f} ms {100 * ms / total:5.1f}%  x{n:4d}  {name[:150]}")
    print("\n## every launch (id, ms, grid, block, kernel)")
    for i, ms, grid, block, name in rows:
        print(f"{i:4d} {ms:9.4f} {grid:>16s} {block:>14s}  {name[:110]}")


if __name__ == "__main__":
    main()
