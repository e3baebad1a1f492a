#!/bin/bash
# attention backward: staggered walk on/off, deterministic mode on/off; CUDA graph re-test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_unet.py tests/test_gpu_network_parity.py tests/test_gpu_query_attention.py -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/f_pytest.log
echo "== stagger on, free-running"; timeout 300 python tools/bench_kernels.py --bwd --batch 16 2>&1 | grep "^{" | tee gpurun_out/f_sites_stagger.jsonl
echo "== stagger on, deterministic"; timeout 300 python tools/bench_kernels.py --bwd --batch 16 --deterministic 2>&1 | grep "^{" | tee gpurun_out/f_sites_stagger_det.jsonl
echo "== stagger off (variant)"; MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_nostagger.so timeout 300 python tools/bench_kernels.py --bwd --batch 16 2>&1 | grep "^{" | tee gpurun_out/f_sites_nostagger.jsonl
echo "== large batch sa6: stagger / det / nostagger"
timeout 300 python tools/bench_kernels.py --bwd --batch 128 --site 0 2>&1 | grep "^{"
timeout 300 python tools/bench_kernels.py --bwd --batch 128 --site 0 --deterministic 2>&1 | grep "^{"
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_nostagger.so timeout 300 python tools/bench_kernels.py --bwd --batch 128 --site 0 2>&1 | grep "^{"
for B in 16 256; do
timeout 600 python bench.py --steps 10 --warmup 4 --batch-per-gpu $B --no-cpu-baseline --cuda-graph > gpurun_out/f_bench_b${B}_graph.json 2> gpurun_out/f_bench_b${B}_graph.err; echo "b$B graph rc $?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/f_bench_*.json")):
    try:
        r=[json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "ms", round(r["ms_per_step"],2), "img/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "host_ms", round(r["host_enqueue_ms_per_step"],2), "launches", r["gpu_launches"], r["config"].get("cuda_graph"), r["clocks"]["sm_mhz"])
    except Exception as e: print(f, "failed", e)
PY
tail -5 gpurun_out/f_bench_b16_graph.err
