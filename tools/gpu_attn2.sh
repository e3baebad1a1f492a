#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -x 2>&1 | tail -3 | tee gpurun_out/attn_test.log
timeout 300 python tools/bench_kernels.py --bwd 2>&1 | tee gpurun_out/kernels_bwd.log
