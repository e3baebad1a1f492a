#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_network_parity.py -m gpu -q -x 2>&1 | tail -3
J() { python - "$1" "$2" <<'PY'
import json, sys
r=[json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
roof=r.get("roofline") or {}
print(sys.argv[1], "ms", round(r["ms_per_step"],2), "value", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "launches", r.get("gpu_launches"), "sm_mhz", (r.get("clocks") or {}).get("sm_mhz"), "roof", roof.get("frac") and round(roof["frac"],3), "loss", r["e2e"].get("last_loss"))
PY
}
for f in 1 0 1 0; do
  MASKUNET_HEAD_STATS=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_headstats$f.json 2>gpurun_out/ab_headstats.err; J headstats$f gpurun_out/ab_headstats$f.json
done
