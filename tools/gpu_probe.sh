#!/bin/bash
mkdir -p gpurun_out
echo "== red.v4" | tee gpurun_out/bwd_dq_red.log
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_red.so timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "backward or full_size" 2>&1 | tail -2 | tee -a gpurun_out/bwd_dq_red.log
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_red.so timeout 300 python tools/bench_kernels.py --batch 16 --bwd 2>&1 | head -3 | tee -a gpurun_out/bwd_dq_red.log
echo "== tma reduce" | tee -a gpurun_out/bwd_dq_red.log
timeout 300 python tools/bench_kernels.py --batch 16 --bwd 2>&1 | head -3 | tee -a gpurun_out/bwd_dq_red.log
