#!/bin/bash
# last check of the final build: full GPU tests, smoke, default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/l_pytest_full.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/l_bench_default.json 2> gpurun_out/l_bench_default.err
python - <<'PY'
import json
r=[json.loads(l) for l in open("gpurun_out/l_bench_default.json") if l.startswith("{")][-1]
print("default ms", round(r["ms_per_step"],2), "value", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "launches", r["gpu_launches"], r["clocks"], "roof", round(r["roofline"]["frac"],3), r["roofline"]["traffic"], "cpu", r["cpu_baseline"]["value"])
PY
