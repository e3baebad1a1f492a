#!/bin/bash
# staged (coalesced) epilogue of the forward projection: tests, then the projection kernels alone for both builds
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_skip_grads.py tests/test_gpu_attention.py tests/test_gpu_network_parity.py tests/test_gpu_unet.py -m gpu -q -x 2>&1 | tail -5
for i in 1 2; do
echo staged; timeout 300 python tools/bench_qkv.py 2>&1 | tail -5
echo direct; MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_p1direct.so timeout 300 python tools/bench_qkv.py 2>&1 | tail -5
done
