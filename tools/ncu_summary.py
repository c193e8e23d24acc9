#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the short per-launch text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/r01_ncu_x.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[start], rows[start + 1]
    print(f"# {path}: ncu --set full --clock-control none (cold cache, serialised; compare shares)")
    for r in rows[start + 2:]:
        d = dict(zip(hdr, r))
        print(f"\n[{d['ID']}] {d['Kernel Name'][:110]}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"    {k:78s} {d[k]:>16s} {units[hdr.index(k)]}")
        stalls = []
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[k]), k))
                except ValueError:
                    pass
        if not stalls:
            for k in hdr:
                if "warp_issue_stalled" in k and k.endswith(".ratio"):
                    try:
                        stalls.append((float(d[k]), k))
                    except ValueError:
                        pass
        for v, k in sorted(stalls, reverse=True)[:6]:
            print(f"    stall {k:72s} {v:16.3f}")


def traffic(paths):
    """--traffic rep...: {kernel base name: DRAM read+write bytes of its LAST profiled launch} as JSON."""
    import json
    import re
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for path in paths:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        hdr, units = rows[start], rows[start + 1]
        for r in rows[start + 2:]:
            d = dict(zip(hdr, r))
            m = re.search(r"(k_[a-z0-9_]+)", d["Kernel Name"])
            if not m:
                continue
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(d[k]) * scale[units[hdr.index(k)]]
            out[m.group(1)] = int(tot)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(sys.argv[2:])
    else:
        main(sys.argv[1])
