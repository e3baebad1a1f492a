#!/bin/bash
# refresh after the tanh GELU: default bench (eager and CUDA graph), BN kernels, launch list + breakdown, and the ncu
# captures of the paired / resident convolution kernels alone (three layer shapes; < 64 MB in all)
mkdir -p gpurun_out
J() { python - "$1" <<'PY'
import json, sys
try:
    r=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    roof=r.get("roofline") or {}
    print(sys.argv[1], "ms", round(r["ms_per_step"],2), "value", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "host_ms", r.get("host_enqueue_ms_per_step") and round(r["host_enqueue_ms_per_step"],1), "sm_mhz", (r.get("clocks") or {}).get("sm_mhz"), "roof", roof.get("frac") and round(roof["frac"],3), "cpu", (r.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/y_bench_default.json 2> gpurun_out/y_bench_default.err; J gpurun_out/y_bench_default.json
timeout 900 python bench.py --steps 10 --warmup 4 --cuda-graph --no-cpu-baseline > gpurun_out/y_bench_b256_graph.json 2> gpurun_out/y_bench_b256_graph.err; J gpurun_out/y_bench_b256_graph.json
timeout 600 python bench.py --steps 5 --warmup 3 --deterministic --no-cpu-baseline > gpurun_out/y_bench_deterministic.json 2> gpurun_out/y_bench_deterministic.err; J gpurun_out/y_bench_deterministic.json
timeout 600 python bench.py --workload city_instance_infer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/y_bench_city.json 2> gpurun_out/y_bench_city.err; J gpurun_out/y_bench_city.json
python tools/with_clocks.py gpurun_out/y_bn_kernels_gbs.json -- python tools/bench_bn.py
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/y_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/y_ncu_launches.log 2>&1
python tools/step_breakdown.py gpurun_out/y_launches.csv 30 | tee gpurun_out/y_step_breakdown.txt | head -8
for sh in "128 128 128" "64 64 128" "256 256 64"; do
  timeout 600 ncu --set full --clock-control none -k regex:conv_fprop_sm100 -s 2 -c 2 -f -o gpurun_out/y_prof_conv_${sh// /_} python tools/run_conv_once.py $sh > gpurun_out/y_ncu_conv.log 2>&1
done
du -sh gpurun_out
