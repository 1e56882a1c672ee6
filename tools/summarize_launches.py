"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            val = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
            rows.append((r["Kernel Name"].split("(")[0], val * scale))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {v:.1f} | {v / tot:.1%} |")
    print(f"| **all** | {len(rows)} | {tot:.1f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1])
