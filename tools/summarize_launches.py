"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share, average.
    python tools/summarize_launches.py launches.csv [--skip N] [--count N] [--title "..."] > summary.txt"""
import argparse
import csv
import re
from collections import OrderedDict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--skip", type=int, default=0)
    ap.add_argument("--count", type=int, default=0)
    ap.add_argument("--title", default="")
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        rows.append((r["Kernel Name"], us))
    rows = rows[a.skip:]
    if a.count:
        rows = rows[:a.count]
    agg = OrderedDict()
    for name, us in rows:
        key = re.sub(r"<.*", "", name).strip()
        key = re.sub(r"\(.*", "", key).strip()
        n, t = agg.get(key, (0, 0.0))
        agg[key] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    if a.title:
        print(a.title)
    print(f"launches in window {len(rows)} total us {total:.1f}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} n={n:4d} total={t:10.1f} us share={100 * t / total:5.1f}% avg={t / n:8.1f}")


if __name__ == "__main__":
    main()
