#!/bin/bash
mkdir -p gpurun_out
echo "== default" | tee gpurun_out/qkv_c256_ab.log
timeout 300 python tools/bench_qkv.py 2>&1 | tail -5 | tee -a gpurun_out/qkv_c256_ab.log
echo "== C=256: NC 64, 1-deep weight ring" | tee -a gpurun_out/qkv_c256_ab.log
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_c256.so timeout 300 python tools/bench_qkv.py 2>&1 | tail -2 | tee -a gpurun_out/qkv_c256_ab.log
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_c256.so timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "projection or module_matches or channels_last" 2>&1 | tail -2 | tee -a gpurun_out/qkv_c256_ab.log
