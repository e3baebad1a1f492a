#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_sm100 -s 1 -c 1 -f -o gpurun_out/prof_attn_bwd_sa6 python tools/bench_kernels.py --batch 16 --bwd --site 0 > gpurun_out/ncu_attn_bwd.log 2>&1
tail -2 gpurun_out/ncu_attn_bwd.log
