#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/attn_test.log
P='import json,sys; d=json.load(sys.stdin); print(d["value"], d["ms_per_step"], d["clocks"], d["roofline"]["achieved"])'
echo "== sleep" | tee gpurun_out/bench_sleep_ab.log
timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "$P" | tee -a gpurun_out/bench_sleep_ab.log
echo "== spin" | tee -a gpurun_out/bench_sleep_ab.log
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_spin.so timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "$P" | tee -a gpurun_out/bench_sleep_ab.log
echo "== sleep again" | tee -a gpurun_out/bench_sleep_ab.log
timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "$P" | tee -a gpurun_out/bench_sleep_ab.log
timeout 300 python tools/bench_kernels.py --batch 16 --bwd --site 0 2>&1 | tail -1 | tee -a gpurun_out/bench_sleep_ab.log
