"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import sys
from collections import defaultdict


def main(path, per_frame=None):
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
    if per_frame:
        # the window is rarely a whole number of frames: per-frame view = mean duration of a launch x its launches per frame
        frame = {k: v / n * per_frame.get(k.split("::")[-1].split("<")[0], 1) for k, (n, v) in agg.items() if "FillFunctor" not in k}
        ftot = sum(frame.values())
        print(f"\n| kernel | launches / frame | us / frame | share of the frame |\n|---|---:|---:|---:|")
        for k, v in sorted(frame.items(), key=lambda kv: -kv[1]):
            print(f"| `{k}` | {per_frame.get(k.split('::')[-1].split('<')[0], 1)} | {v:.1f} | {v / ftot:.1%} |")
        print(f"| **frame** | | {ftot:.1f} | 100% |")


if __name__ == "__main__":
    # optional second argument: kernel=launches-per-frame pairs, e.g. chain_tc_kernel=4
    pf = dict((a.split("=")[0], int(a.split("=")[1])) for a in sys.argv[2:])
    main(sys.argv[1], pf or None)
