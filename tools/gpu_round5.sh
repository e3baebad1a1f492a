#!/bin/bash
mkdir -p gpurun_out
# attention kernels of the first step at the bench configuration (backward runs the 16384-token site first)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:attn_.wd_sm100 -c 12 -f -o gpurun_out/prof_bench_attn python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench_attn.log 2>&1
tail -2 gpurun_out/ncu_bench_attn.log | cut -c1-300
timeout 1500 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-300
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-300
timeout 600 python tools/bench_kernels.py --bwd --sweep --out gpurun_out/kernel_sweep.json 2>&1 | tail -20
