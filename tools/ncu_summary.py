#!/usr/bin/env python
"""Summarises an `ncu --set full` report into profiles/<name>.md + .json: per kernel (launches of the same name and grid
averaged) the duration, DRAM traffic, L2 reduction / atomic sectors, shared-memory atomic bank conflicts, issue-slot use and
occupancy -- the counters VERDICT.md asks for next to every non-tensor kernel.

    python tools/ncu_summary.py gpurun_out/r2a_vote_shot.ncu-rep profiles/r02_vote_shot_ncu_full "command line that made it"
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "duration_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "l2_red_sectors"),
    ("lts__t_sectors_srcunit_tex_op_atom.sum", "l2_atom_sectors"),
    ("lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "l2_atomic_unit_pct"),
    ("smsp__inst_executed_op_shared_atom.sum", "smem_atom_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "smem_atom_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smem_atom_wavefronts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])


def main():
    rep, out = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    groups = OrderedDict()
    for r in data:
        name = r[col["Kernel Name"]]
        key = (name, r[col["launch__grid_size"]] if "launch__grid_size" in col else "")
        groups.setdefault(key, []).append(r)
    summary = []
    for (name, grid), rs in groups.items():
        item = {"kernel": name, "launches": len(rs)}
        for m, short in METRICS.items():
            if m not in col:
                continue
            vals = []
            for r in rs:
                try:
                    vals.append(float(r[col[m]].replace(",", "")))
                except ValueError:
                    pass
            if not vals:
                continue
            v = sum(vals) / len(vals)
            u = units[col[m]]
            if short == "duration_us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            if short.endswith("_MB"):
                v = v / 1e6 if u in ("byte", "B") else (v / 1e3 if u in ("Kbyte", "KB") else (v * 1e3 if u in ("Gbyte", "GB") else v))
            item[short] = round(v, 4)
        summary.append(item)
    json.dump({"report": rep, "command": cmd, "kernels": summary}, open(out + ".json", "w"), indent=1)
    keys = [k for k in METRICS.values() if any(k in it for it in summary)]
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary of `{rep}`\n\nCommand: `{cmd}`\n\nPer-launch averages over the launches of each (kernel, grid); "
                "numbers under the profiler are never bench values (cold caches, serialised, clocks not locked).\n\n")
        f.write("| kernel | n | " + " | ".join(keys) + " |\n|---|---:|" + "---:|" * len(keys) + "\n")
        for it in summary:
            f.write(f"| `{it['kernel'][:70]}` | {it['launches']} | " + " | ".join(str(it.get(k, "")) for k in keys) + " |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
