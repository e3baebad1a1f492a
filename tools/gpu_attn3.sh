#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_unet.py -q -x 2>&1 | tail -3 | tee gpurun_out/attn_test.log
echo "== poly every 4"; timeout 300 python tools/bench_kernels.py 2>&1 | tee gpurun_out/kernels_fwd_poly4.log
echo "== no poly"; MASKUNET_B200_LIB=$PWD/maskunet_b200/libmaskunet_b200_nopoly.so timeout 300 python tools/bench_kernels.py 2>&1 | tee gpurun_out/kernels_fwd_nopoly.log
timeout 1500 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_default.json
