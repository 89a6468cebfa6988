#!/usr/bin/env python
"""Opcode histogram and hottest SASS lines (warp-stall samples) of the first kernel in an ncu report.

    python profiles/sass_hotspots.py gpurun_out/prof.ncu-rep [n_lines]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main(path, n=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[hdr_idx[0]]
    body = rows[hdr_idx[0] + 1:(hdr_idx[1] - 1 if len(hdr_idx) > 1 else None)]
    si, ei, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    body = [r for r in body if len(r) > ei and r[ei].isdigit()]
    warps = max(int(r[ei]) for r in body[:4]) or 1
    ex, sm = Counter(), Counter()
    for r in body:
        t = r[src].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ex[op] += int(r[ei])
        sm[op] += int(r[si])
    print("kernel:", rows[0][1][:100])
    print("samples:", sum(sm.values()), " warps:", warps, " instr/warp:", sum(ex.values()) // warps)
    print("executed per warp:", [(k, v // warps) for k, v in ex.most_common(14)])
    print("stall samples by opcode:", sm.most_common(12))
    for r in sorted(body, key=lambda r: -int(r[si]))[:n]:
        print("%6s %9s  %s" % (r[si], r[ei], r[src][:100]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
