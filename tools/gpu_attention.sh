#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_query_attention.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/attn_test.log
timeout 600 python tools/bench_kernels.py --batch 16 --bwd 2>&1 | tee gpurun_out/kernels_bwd.log
