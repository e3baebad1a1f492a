#!/bin/bash
# split-range upsample + concat kernels: tests, then an A/B of the default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_unet.py tests/test_gpu_network_parity.py -m gpu -q -x 2>&1 | tail -4
J() { python - "$1" "$2" <<'PY'
import json, sys
r=[json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
roof=r.get("roofline") or {}
print(sys.argv[1], "ms", round(r["ms_per_step"],2), "value", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "launches", r.get("gpu_launches"), "sm_mhz", (r.get("clocks") or {}).get("sm_mhz"), "roof", roof.get("frac") and round(roof["frac"],3), roof.get("kernel_ms") and round(roof["kernel_ms"],2), "loss", r["e2e"].get("last_loss"))
PY
}
for i in 1 2; do
  for f in 1 0; do
    MU_UPCAT_SPLIT=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_upcat${f}_$i.json 2>gpurun_out/ab_upcat.err; J upcat_split$f gpurun_out/ab_upcat${f}_$i.json
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upcat -c 24 --csv --log-file gpurun_out/x_upcat_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/x_upcat_launches.csv')):
    if len(r) > 10 and r[0] != 'ID' and 'upcat' in r[4]: print(r[4][:44], r[8], float(r[-1]) / 1e3, 'us')
PY
