"""Per-CUDA-line stall samples of an ncu report (needs -lineinfo and --import-source on).

    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv [--kernel-name regex:...] > src.csv
    python tools/ncu_lines.py src.csv [launch_index] [top_n]

The CSV is a sequence of (file, kernel launch) blocks; the rows whose first column is a line number carry the
metrics aggregated over that line's SASS.
"""
import csv
import sys
from collections import defaultdict


def main(path, launch=None, top=40):
    rows = list(csv.reader(open(path, newline="")))
    blocks, cur, fname, hdr = [], None, None, None
    launches = defaultdict(list)   # function block order -> list of (file, line, src, samples, stall dict)
    seen_files = defaultdict(int)
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1]
            seen_files[fname] += 1
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                n = int(d.get("# Samples", "0") or 0)
            except ValueError:
                n = 0
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
            launches[seen_files[fname] - 1].append((fname.split("/")[-1], int(r[0]), r[1], n, stalls))
    for li in sorted(launches):
        if launch is not None and li != launch:
            continue
        data = launches[li]
        tot = sum(x[3] for x in data) or 1
        print(f"== launch {li}: {tot} samples")
        for f, ln, src, n, st in sorted(data, key=lambda x: -x[3])[:top]:
            why = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(f"{100 * n / tot:5.1f}%  {f}:{ln:<4d} {src.strip()[:90]:90s} | {why}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
