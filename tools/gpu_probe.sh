#!/bin/bash
mkdir -p gpurun_out
echo "== default (acc1 @64, NC128 @128)" | tee gpurun_out/qkv_nc_ab.log
timeout 300 python tools/bench_qkv.py 2>&1 | tail -3 | tee -a gpurun_out/qkv_nc_ab.log
echo "== NC 64 @ C=128" | tee -a gpurun_out/qkv_nc_ab.log
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_nc64.so timeout 300 python tools/bench_qkv.py 2>&1 | tail -1 | tee -a gpurun_out/qkv_nc_ab.log
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_nc64.so timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "projection or module_matches or channels_last" 2>&1 | tail -2 | tee -a gpurun_out/qkv_nc_ab.log
timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "projection or module_matches or channels_last" 2>&1 | tail -2 | tee -a gpurun_out/qkv_nc_ab.log
