#!/bin/bash
# round 2, final 1-GPU pass (statistics fold in the attention backward, skip-gradient folding, staged projection epilogues,
# fused LayerNorm re-view, split upsample kernels): full GPU tests, smoke, both bench arms, other workloads, clocked
# per-kernel numbers, ncu launch list + step breakdown, ncu --set full of the attention kernels inside the benchmark step
mkdir -p gpurun_out
J() { python - "$1" <<'PY'
import json, sys
try:
    r=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    roof=r.get("roofline") or {}
    print(sys.argv[1], "ms", round(r["ms_per_step"],2), "value", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "host_ms", r.get("host_enqueue_ms_per_step") and round(r["host_enqueue_ms_per_step"],1), "launches", r.get("gpu_launches"), "sm_mhz", (r.get("clocks") or {}).get("sm_mhz"), "roof", roof.get("kernel"), roof.get("frac") and round(roof["frac"],3), "cpu", (r.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/f_pytest_full.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err; J gpurun_out/f_bench_default.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; J gpurun_out/f_bench_reference.json
timeout 900 python bench.py --steps 10 --warmup 4 --cuda-graph --no-cpu-baseline > gpurun_out/f_bench_b256_graph.json 2> gpurun_out/f_bench_b256_graph.err; J gpurun_out/f_bench_b256_graph.json
timeout 600 python bench.py --steps 5 --warmup 3 --deterministic --no-cpu-baseline > gpurun_out/f_bench_deterministic.json 2> gpurun_out/f_bench_deterministic.err; J gpurun_out/f_bench_deterministic.json
timeout 600 python bench.py --workload coco_panoptic --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_bench_coco.json 2> gpurun_out/f_bench_coco.err; J gpurun_out/f_bench_coco.json
timeout 600 python bench.py --workload city_instance_infer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_city.json 2> gpurun_out/f_bench_city.err; J gpurun_out/f_bench_city.json
python tools/with_clocks.py gpurun_out/f_attn_kernels_final.json -- python tools/bench_kernels.py --bwd --batch 64
python tools/with_clocks.py gpurun_out/f_qkv_project.json -- python tools/bench_qkv.py
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/f_ncu_launches.log 2>&1
python tools/step_breakdown.py gpurun_out/f_launches.csv 30 | tee gpurun_out/f_step_breakdown.txt | head -12
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:attn_.wd_sm100 -c 12 -f -o gpurun_out/f_prof_attn python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/f_ncu_attn.log 2>&1
tail -2 gpurun_out/f_ncu_attn.log
du -sh gpurun_out
