#!/bin/bash
# A/B of PAR kernel variants on one B200: parity subset, then a short bench, per configuration.
# Each argument is a comma-separated list of environment settings, e.g.
#   gpurun -- 'bash profiles/ab_par.sh COSA_PAR_STEP=prop3 COSA_PAR_STEP=prop3,COSA_PAR_COOP=0 COSA_PAR_STEP=smem'
for cfg in "$@"; do
  echo "== $cfg"
  tag=$(echo "$cfg" | tr -c 'A-Za-z0-9\n' '_')
  envs=$(echo "$cfg" | tr ',' ' ')
  env $envs timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q -m gpu -k "par or cam2mask or full_size" 2>&1 | tail -2
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$tag.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_$tag.log") if x.startswith("{")]
if not l:
    print(open("gpurun_out/bench_$tag.log").read()[-1500:])
else:
    d=json.loads(l[-1])
    print(round(d["value"]), round(d["ms_per_step"],3), [(k["kernel"],k["ms_per_step"]) for k in d["kernels"][:3]])
PY
done
