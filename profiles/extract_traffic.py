#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch for every kernel of an `ncu --set full` report
-> profiles/dram_traffic.json (read by bench.py for roofline.traffic; the file names the configuration the capture was
made on, and bench.py reports the traffic only when it runs that configuration).

    python profiles/extract_traffic.py <config key, e.g. voc_b32_448x448> gpurun_out/prof.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(config, paths):
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dram_traffic.json")
    kernels, sources = {}, {}
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
            kernels[name] = int(sum(vals) / len(vals))
        sources[os.path.basename(path)] = sorted(acc)
    out = {"config": config, "kernels": kernels, "sources": sources,
           "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches)"}
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
