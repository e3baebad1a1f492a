#!/bin/bash
# usage: gpu_ab.sh name1=path1.so name2=path2.so ...   ("default" = the in-tree library): alternating bench.py runs on one box
mkdir -p gpurun_out
J() { python - "$1" "$2" <<'PY'
import json, sys
r=[json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
roof=r.get("roofline") or {}
print(sys.argv[1], "ms", round(r["ms_per_step"],2), "value", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "sm_mhz", (r.get("clocks") or {}).get("sm_mhz"), "roof", roof.get("frac") and round(roof["frac"],3), roof.get("kernel_ms") and round(roof["kernel_ms"],2))
PY
}
for i in 1 2; do
for spec in "$@"; do
  name=${spec%%=*}; lib=${spec#*=}
  if [ "$lib" = "default" ]; then unset MASKUNET_B200_LIB; else export MASKUNET_B200_LIB=$PWD/$lib; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${name}_$i.json 2>/dev/null; J $name gpurun_out/ab_${name}_$i.json
done; done
