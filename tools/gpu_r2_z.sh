#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_query_attention.py -m gpu -q -x 2>&1 | tail -6
echo fold; timeout 300 python tools/bench_kernels.py --bwd --batch 64 --site 0 2>&1 | tail -1
timeout 300 python tools/bench_kernels.py --bwd --batch 64 --site 1 2>&1 | tail -1
echo nofold; MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_nofold.so timeout 300 python tools/bench_kernels.py --bwd --batch 64 --site 0 2>&1 | tail -1
MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_nofold.so timeout 300 python tools/bench_kernels.py --bwd --batch 64 --site 1 2>&1 | tail -1
