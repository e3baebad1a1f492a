#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "tcgen05 or module_matches or sdpa or channels_last" 2>&1 | tail -5 | tee gpurun_out/attn_test.log
timeout 600 python tools/bench_kernels.py --batch 16 --bwd 2>&1 | tee gpurun_out/kernels_bwd.log
