#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rXX_bench_launches_summary.txt
"""
import collections
import csv
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], v_us))
    tot = sum(v for _, v in rows)
    agg = collections.defaultdict(lambda: [0.0, 0])
    for k, v in rows:
        agg[k][0] += v
        agg[k][1] += 1
    print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 python bench.py --steps 2 --warmup 3 --no-cpu")
    print("# (cold-cache, serialised launch times: compare SHARES; first 3000 launches of the run)")
    print(f"total {tot:.1f} us over {len(rows)} launches")
    for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{v:10.1f} us {100 * v / tot:5.1f}% n={n:5d} avg {v / n:8.1f} us  {k[:90]}")


if __name__ == "__main__":
    main(sys.argv[1])
