#!/bin/bash
echo nopipe; MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_bnopipe.so timeout 300 python tools/bench_kernels.py --bwd --batch 64 2>&1 | tail -6 | cut -c60-200
echo pipe; timeout 300 python tools/bench_kernels.py --bwd --batch 64 2>&1 | tail -6 | cut -c60-200
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_query_attention.py -m gpu -q -x 2>&1 | tail -3
