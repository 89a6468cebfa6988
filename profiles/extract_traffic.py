#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch for every kernel of an `ncu --set full` report
-> profiles/dram_traffic.json (read by bench.py for roofline.traffic).

    python profiles/extract_traffic.py gpurun_out/prof.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(paths):
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dram_traffic.json")
    traffic = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        acc = {}
        for r in data:
            name = r[idx["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").replace("cosa::", "").strip()
            rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT[units[idx["dram__bytes_write.sum"]]]
            acc.setdefault(name, []).append(rd + wr)
        for name, vals in acc.items():
            traffic[name] = int(sum(vals) / len(vals))
        traffic.setdefault("_source", {})[os.path.basename(path)] = sorted(acc)
    json.dump(traffic, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1:])
