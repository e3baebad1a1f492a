#!/bin/bash
mkdir -p gpurun_out
J() { python - "$1" <<'PY'
import json, sys
r=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
roof=r.get("roofline") or {}
print(sys.argv[1], "ms", round(r["ms_per_step"],2), "value", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "sm_mhz", (r.get("clocks") or {}).get("sm_mhz"), "roof", roof.get("frac") and round(roof["frac"],3), roof.get("kernel_ms"))
PY
}
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/u_new$i.json 2>/dev/null; J gpurun_out/u_new$i.json
MU_CONV_PAIR=0 MU_CONV_RES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/u_old$i.json 2>/dev/null; J gpurun_out/u_old$i.json
MU_CONV_PAIR=1 MU_CONV_RES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/u_pair$i.json 2>/dev/null; J gpurun_out/u_pair$i.json
done
