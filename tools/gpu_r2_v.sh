#!/bin/bash
# round 2, 1-GPU pass after the bf16 dQ reduce / CTA-pair convolution / cross-entropy changes: tests, smoke, benches of
# every workload, clocked per-kernel numbers, ncu launch list + step breakdown, ncu --set full of the attention kernels
# and of the paired convolution inside the benchmark step
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
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/v_bench_default.json 2> gpurun_out/v_bench_default.err; J gpurun_out/v_bench_default.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/v_bench_reference.json 2> gpurun_out/v_bench_reference.err; J gpurun_out/v_bench_reference.json
timeout 600 python bench.py --steps 5 --warmup 3 --deterministic --no-cpu-baseline > gpurun_out/v_bench_deterministic.json 2> gpurun_out/v_bench_deterministic.err; J gpurun_out/v_bench_deterministic.json
timeout 600 python bench.py --workload coco_panoptic --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/v_bench_coco.json 2> gpurun_out/v_bench_coco.err; J gpurun_out/v_bench_coco.json
timeout 600 python bench.py --workload city_instance_infer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_city.json 2> gpurun_out/v_bench_city.err; J gpurun_out/v_bench_city.json
timeout 600 python bench.py --steps 20 --warmup 4 --batch-per-gpu 64 --no-cpu-baseline --cuda-graph > gpurun_out/v_bench_b64_graph.json 2> gpurun_out/v_bench_b64_graph.err; J gpurun_out/v_bench_b64_graph.json
timeout 900 python bench.py --workload kernel_sweep --full-sweep > gpurun_out/v_kernel_sweep.json 2> gpurun_out/v_kernel_sweep.err; echo "sweep rc $?"
python tools/with_clocks.py gpurun_out/v_attn_kernels_final.json -- python tools/bench_kernels.py --bwd --batch 64
python tools/with_clocks.py gpurun_out/v_bn_kernels_gbs.json -- python tools/bench_bn.py
python tools/with_clocks.py gpurun_out/v_conv_layers_ours.json -- python tools/bench_conv_ours.py
python tools/with_clocks.py gpurun_out/v_qkv_project.json -- python tools/bench_qkv.py
python tools/with_clocks.py gpurun_out/v_ce_kernel.json -- python tools/bench_ce.py
python tools/conv_totals.py gpurun_out/v_conv_layers_ours.json profiles/r02_conv_layers_library.json 2>&1 | tail -3
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/v_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/v_ncu_launches.log 2>&1
python tools/step_breakdown.py gpurun_out/v_launches.csv 30 | tee gpurun_out/v_step_breakdown.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:attn_.wd_sm100 -c 12 -f -o gpurun_out/v_prof_attn python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/v_ncu_attn.log 2>&1
tail -2 gpurun_out/v_ncu_attn.log
# paired convolution kernels alone: 128 -> 128 and 64 -> 64 at 128 x 128, 256 -> 256 at 64 x 64 (forward + data gradient)
for sh in "128 128 128" "64 64 128" "256 256 64"; do
  timeout 600 ncu --set full --clock-control none -k regex:conv_fprop_sm100 -s 2 -c 2 -f -o gpurun_out/v_prof_conv_${sh// /_} python tools/run_conv_once.py $sh > gpurun_out/v_ncu_conv.log 2>&1
done
tail -2 gpurun_out/v_ncu_conv.log
ls -la gpurun_out/
