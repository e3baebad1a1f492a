#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_.wd_sm100 -s 4 -c 2 -f -o gpurun_out/prof_attn_sa6 python tools/bench_kernels.py --batch 16 --bwd --site 0 > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
ls -la gpurun_out/
