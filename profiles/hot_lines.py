#!/usr/bin/env python
"""Hottest SASS lines (warp-stall samples + dominant stall reason) of one kernel in an ncu report.

    python profiles/hot_lines.py gpurun_out/prof.ncu-rep <kernel regex> [n_lines]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main(path, pat, n=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass",
                          "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[hdr_idx[0]]
    body = rows[hdr_idx[0] + 1:(hdr_idx[1] - 1 if len(hdr_idx) > 1 else None)]
    si, ei, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    stalls = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    body = [r for r in body if len(r) > ei and r[ei].isdigit()]
    warps = max(int(r[ei]) for r in body[:4]) or 1
    tot = Counter()
    for r in body:
        for i, c in stalls:
            tot[c] += int(r[i] or 0)
    print("kernel:", rows[0][1][:110])
    print("samples:", sum(int(r[si]) for r in body), " warps:", warps, " instr/warp:", sum(int(r[ei]) for r in body) // warps)
    print("stall totals:", tot.most_common(8))
    order = sorted(range(len(body)), key=lambda k: -int(body[k][si]))[:n]
    for k in sorted(order):
        r = body[k]
        top = max(stalls, key=lambda ic: int(r[ic[0]] or 0))[1]
        print("%4d %6s %9s %-14s %s" % (k, r[si], r[ei], top[6:], r[src][:90]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
