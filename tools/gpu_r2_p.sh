#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_$i.json 2>/dev/null
python - gpurun_out/p_$i.json <<'PY'
import json, sys
r=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
print("ms", round(r["ms_per_step"],2), "e2e ms", round(r["e2e"]["ms_per_step"],2), "steps", r["e2e"]["ms_steps_host_wall"], "clocks", r["clocks"]["sm_mhz"], r["e2e"]["clocks"])
PY
done
