#!/usr/bin/env python
"""Print the headline and the per-kernel table of a bench.py JSON line (file argument or stdin)."""
import json
import sys

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
line = [l for l in src if l.startswith("{")][-1]
d = json.loads(line)
print("%.0f %s  %.3f ms/step  e2e=%s  launches=%s" % (d["value"], d["unit"], d["ms_per_step"],
      (round(d["e2e"]["value"]) if d.get("e2e") else None), d.get("gpu_launches")))
for k in d.get("kernels", []):
    print("  %-34s x%-4g %8.4f ms/step  frac=%s" % (k["kernel"], k["launches_per_step"], k["ms_per_step"], k["frac"]))
if d.get("cpu_baseline"):
    print("  cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
