#!/usr/bin/env python
"""Summarise an Nsight Compute report (one block per distinct kernel, first captured launch) into text.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_full_summary.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    print("# %s - first captured launch of each kernel (ncu --set full --clock-control none)" % path)
    for r in data:
        name = r[idx["Kernel Name"]]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:
            continue
        print("\n== %s" % name)
        for m in METRICS:
            if m in idx:
                print("   %-66s %14s %s" % (m, r[idx[m]], units[idx[m]]))
        if "dram__bytes_read.sum" in idx:
            unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd = float(r[idx["dram__bytes_read.sum"]]) * unit.get(units[idx["dram__bytes_read.sum"]], 1)
            wr = float(r[idx["dram__bytes_write.sum"]]) * unit.get(units[idx["dram__bytes_write.sum"]], 1)
            print("   %-66s %14.0f byte" % ("traffic = dram read + write", rd + wr))
    print("\n# launches captured per kernel: %s" % seen)


if __name__ == "__main__":
    main(sys.argv[1])
