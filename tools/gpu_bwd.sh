#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "tcgen05_backward" 2>&1 | tail -30 | tee gpurun_out/bwd_test.log
timeout 600 python -m pytest tests -m gpu -q -x -k "module_matches and bfloat16" 2>&1 | tail -30 | tee -a gpurun_out/bwd_test.log
timeout 600 python tools/bench_kernels.py --batch 16 --bwd 2>&1 | tee gpurun_out/kernels_bwd.log
