#!/usr/bin/env python
"""Pretty-print the JSON line bench.py wrote (per-kernel table): python profiles/show_bench.py gpurun_out/bench.log"""
import json
import sys

line = [l for l in open(sys.argv[1]) if l.startswith("{")][-1]
d = json.loads(line)
for k in ("value", "ms_per_step", "gpu_launches", "clocks", "e2e", "cpu_baseline", "roofline", "loss"):
    print(k, d.get(k))
print(d["config"])
tot = sum(k["ms_per_step"] for k in d["kernels"])
print("sum of kernel ms/step %.4f" % tot)
for k in d["kernels"]:
    print("%-34s n/step=%5.1f avg_ms=%8.4f ms/step=%8.4f share=%5.1f%% GB/s=%s frac=%s" % (
        k["kernel"], k["launches_per_step"], k["avg_ms"], k["ms_per_step"], 100 * k["ms_per_step"] / tot,
        k["achieved_gbs"], k["frac"]))
