#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_query_attention.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/attn_test.log
echo "== speculative" | tee gpurun_out/kernels_fwd_spec.log
timeout 600 python tools/bench_kernels.py --batch 16 2>&1 | tee -a gpurun_out/kernels_fwd_spec.log
echo "== max first" | tee -a gpurun_out/kernels_fwd_spec.log
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_nospec.so timeout 600 python tools/bench_kernels.py --batch 16 2>&1 | tee -a gpurun_out/kernels_fwd_spec.log
