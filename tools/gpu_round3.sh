#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_full.log
timeout 1500 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_default.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_.wd_sm100 -s 4 -c 2 -f -o gpurun_out/prof_attn_sa6_r3 python tools/bench_kernels.py --batch 16 --bwd --site 0 > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
